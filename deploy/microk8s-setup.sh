#!/bin/bash
# Single-machine dev cluster: MicroK8s + GPU addon + local registry + the chart.
set -e
sudo snap install microk8s --classic
sudo microk8s enable dns storage registry gpu helm3
sudo microk8s helm3 install adaptdl-b200 "$(dirname "$0")/helm/adaptdl-b200-sched" \
    --namespace adaptdl --create-namespace --set registry.enabled=false
