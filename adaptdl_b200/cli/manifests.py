"""Manifest construction for the CLI (pure functions; unit-tested).

* :func:`prepare_job` -- turn the user's ``adaptdljob.yaml`` into the object
  that is created: image digest + args injected, checkpoint/share volume
  (an RWX PVC mounted at ``/adaptdl/checkpoint`` and ``/adaptdl/share`` with
  ``ADAPTDL_CHECKPOINT_PATH`` / ``ADAPTDL_SHARE_PATH``), optional TensorBoard
  volume (``ADAPTDL_TENSORBOARD_LOGDIR``), image pull secret.
* PVC / copy-pod / TensorBoard deployment+service manifests.
"""

import copy
import os
import uuid

TENSORBOARD_PREFIX = "adaptdl-tensorboard-"
PVC_VOLUME = "adaptdl-pvc"
CHECKPOINT_MOUNT = "/adaptdl/checkpoint"
SHARE_MOUNT = "/adaptdl/share"
TENSORBOARD_MOUNT = "/adaptdl/tensorboard"


def prepare_job(resource, image, args, name=None, pull_secret=None,
                tensorboard=None, pvc_name=None):
    """Returns ``(job_object, pvc_name)``."""
    job = copy.deepcopy(resource)
    pod_spec = job["spec"]["template"]["spec"]
    main = pod_spec["containers"][0]
    main["image"] = image
    main["args"] = list(args)
    if pull_secret:
        pod_spec["imagePullSecrets"] = [{"name": pull_secret}]
    if name is not None:
        job["metadata"].pop("name", None)
        job["metadata"]["generateName"] = name + "-"
    volumes = pod_spec.setdefault("volumes", [])
    if tensorboard is not None:
        volumes.append({
            "name": "adaptdl-tensorboard",
            "persistentVolumeClaim": {
                "claimName": TENSORBOARD_PREFIX + tensorboard}})
        for container in pod_spec["containers"]:
            container.setdefault("volumeMounts", []).append(
                {"name": "adaptdl-tensorboard",
                 "mountPath": TENSORBOARD_MOUNT})
            container.setdefault("env", []).append(
                {"name": "ADAPTDL_TENSORBOARD_LOGDIR",
                 "value": TENSORBOARD_MOUNT})
    pvc_name = pvc_name or "adaptdl-pvc-{}".format(uuid.uuid4())
    for container in pod_spec["containers"]:
        mounts = container.setdefault("volumeMounts", [])
        mounts.append({"name": PVC_VOLUME, "mountPath": CHECKPOINT_MOUNT,
                       "subPath": "adaptdl/checkpoint"})
        mounts.append({"name": PVC_VOLUME, "mountPath": SHARE_MOUNT,
                       "subPath": "adaptdl/share"})
        container.setdefault("env", []).extend([
            {"name": "ADAPTDL_CHECKPOINT_PATH", "value": CHECKPOINT_MOUNT},
            {"name": "ADAPTDL_SHARE_PATH", "value": SHARE_MOUNT}])
    volumes.append({"name": PVC_VOLUME,
                    "persistentVolumeClaim": {"claimName": pvc_name}})
    return job, pvc_name


LAUNCHER = ["-m", "adaptdl_b200.launch"]


def use_node_pods(job):
    """Turn a one-GPU-per-pod job into a one-pod-per-node job (in place):
    ``spec.podPerNode`` plus the replica launcher in front of every
    container command that starts a Python interpreter
    (``python train.py`` -> ``python -m adaptdl_b200.launch train.py``).
    Containers with another entry point are left alone -- they must start
    ``ADAPTDL_LOCAL_REPLICAS`` processes themselves."""
    job["spec"]["podPerNode"] = True
    for container in job["spec"]["template"]["spec"]["containers"]:
        command = container.get("command") or []
        if command and os.path.basename(command[0]).startswith("python") \
                and command[1:3] != LAUNCHER:
            container["command"] = [command[0]] + LAUNCHER + command[1:]
    return job


def choose_storageclass(storage_classes, name=None):
    """Pick the StorageClass for the RWX checkpoint volume: an explicit
    ``name``, else the first known RWX-capable provisioner (hostpath,
    CephFS, EFS), else the cluster default."""
    by_name = {sc["metadata"]["name"]: sc for sc in storage_classes}
    if name is not None:
        if name not in by_name:
            raise SystemExit("Error: StorageClass {} not found".format(name))
        return name
    preferred = ("microk8s.io/hostpath", "ceph.rook.io/block",
                 "rook-ceph.cephfs.csi.ceph.com", "efs.csi.aws.com",
                 "example.com/aws-efs")
    for sc in storage_classes:
        if sc.get("provisioner") in preferred:
            return sc["metadata"]["name"]
    for sc in storage_classes:
        ann = sc["metadata"].get("annotations") or {}
        if ann.get("storageclass.kubernetes.io/is-default-class") == "true":
            return sc["metadata"]["name"]
    raise SystemExit("Error: no suitable StorageClass found; pass "
                     "--checkpoint-storage-class")


def pvc_manifest(name, storage_class, size, owner_metadata=None):
    meta = {"name": name}
    if owner_metadata:
        meta["ownerReferences"] = [{
            "apiVersion": "adaptdl.petuum.com/v1", "kind": "AdaptDLJob",
            "name": owner_metadata["name"], "uid": owner_metadata["uid"]}]
    return {"apiVersion": "v1", "kind": "PersistentVolumeClaim",
            "metadata": meta,
            "spec": {"accessModes": ["ReadWriteMany"],
                     "resources": {"requests": {"storage": size}},
                     "storageClassName": storage_class}}


def copy_pod_manifest(pvc_name, uid):
    """A sleeping alpine pod with the job's PVC mounted at /adaptdl_pvc, the
    source of ``kubectl cp``."""
    return {"apiVersion": "v1", "kind": "Pod",
            "metadata": {"name": "adaptdl-copy-{}".format(uid)},
            "spec": {
                "containers": [{
                    "name": "copy", "image": "alpine:latest",
                    "command": ["sleep", "1000000"],
                    "volumeMounts": [{"name": PVC_VOLUME,
                                      "mountPath": "/adaptdl_pvc",
                                      "subPath": "adaptdl"}]}],
                "volumes": [{"name": PVC_VOLUME,
                             "persistentVolumeClaim": {
                                 "claimName": pvc_name}}],
                "restartPolicy": "Never"}}


def tensorboard_manifests(name, storage_class, size="1Gi",
                          image="tensorflow/tensorflow:latest",
                          nodeport=False):
    """Deployment + Service + PVC of a named TensorBoard instance
    (``nodeport``: expose the service on a node port instead of ClusterIP,
    reference ``adaptdl tensorboard create --nodeport``)."""
    full = TENSORBOARD_PREFIX + name
    labels = {"app": "adaptdl-tensorboard", "adaptdl/tensorboard": name}
    deployment = {
        "apiVersion": "apps/v1", "kind": "Deployment",
        "metadata": {"name": full, "labels": labels},
        "spec": {"replicas": 1, "selector": {"matchLabels": labels},
                 "template": {"metadata": {"labels": labels}, "spec": {
                     "containers": [{
                         "name": "tensorboard", "image": image,
                         "command": ["tensorboard", "--logdir",
                                     TENSORBOARD_MOUNT, "--host", "0.0.0.0",
                                     "--port", "6006"],
                         "ports": [{"containerPort": 6006}],
                         "volumeMounts": [{"name": "logs",
                                           "mountPath": TENSORBOARD_MOUNT}]}],
                     "volumes": [{"name": "logs", "persistentVolumeClaim": {
                         "claimName": full}}]}}}}
    service = {"apiVersion": "v1", "kind": "Service",
               "metadata": {"name": full, "labels": labels},
               "spec": {"selector": labels,
                        "ports": [{"name": "http", "port": 6006,
                                   "targetPort": 6006}]}}
    if nodeport:
        service["spec"]["type"] = "NodePort"
    return [pvc_manifest(full, storage_class, size), deployment, service]


def summarize_jobs(items, now):
    """Rows for ``ls``: name, status, start, runtime, replicas, restarts."""
    from datetime import datetime
    rows = []
    for job in items:
        status = job.get("status") or {}
        start = datetime.strptime(job["metadata"]["creationTimestamp"],
                                  "%Y-%m-%dT%H:%M:%SZ")
        end = status.get("completionTimestamp")
        if end:
            end = datetime.fromisoformat(end.replace("Z", "+00:00")) \
                .replace(tzinfo=None)
        else:
            end = now
        rows.append({"name": job["metadata"]["name"],
                     "phase": status.get("phase", "Pending"),
                     "start_time": start,
                     "run_time": str(end - start).split(".")[0],
                     "replicas": status.get("replicas") or "N/A",
                     "restarts": status.get("group", 0)})
    return rows
