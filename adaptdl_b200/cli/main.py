"""``adaptdl-b200`` command line (see package docstring)."""

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import uuid
from datetime import datetime

from adaptdl_b200.cli import manifests


def _kubectl(*args, input_obj=None, capture=True):
    cmd = ["kubectl"] + list(args)
    data = json.dumps(input_obj).encode() if input_obj is not None else None
    try:
        if capture:
            return subprocess.check_output(cmd, input=data).decode()
        subprocess.check_call(cmd)
        return ""
    except FileNotFoundError:
        raise SystemExit("Error: kubectl not found on PATH")
    except subprocess.CalledProcessError as exc:
        raise SystemExit("Error: {} failed ({})".format(" ".join(cmd),
                                                        exc.returncode))


def _create(obj):
    return json.loads(_kubectl("create", "-f", "-", "-o", "json",
                               input_obj=obj))


def _load_yaml(path):
    """Job manifest from a file, or from stdin for ``-f -``."""
    import yaml
    if path == "-":
        return yaml.safe_load(sys.stdin)
    with open(path) as f:
        return yaml.safe_load(f)


class _RegistryTunnel(object):
    """``localhost:port`` -> the in-cluster insecure registry, through the
    API server's service proxy (``cli/proxy.py``, what the reference does
    with mitmproxy); ``kubectl port-forward`` if the kubeconfig cannot be
    used that way (exec credential plugins)."""

    def __init__(self, port, namespace="default"):
        self._stack = None
        self._forward = None
        try:
            import contextlib
            from adaptdl_b200.cli.proxy import service_proxy
            self._stack = contextlib.ExitStack()
            self._stack.enter_context(service_proxy(
                namespace, "adaptdl-registry:registry", listen_port=port))
        except Exception:  # noqa: BLE001 - any failure: plain port-forward
            self._stack = None
            self._forward = subprocess.Popen(
                ["kubectl", "port-forward", "service/adaptdl-registry",
                 "{}:5000".format(port)],
                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            time.sleep(2)

    def terminate(self):
        if self._stack is not None:
            self._stack.close()
        if self._forward is not None:
            self._forward.terminate()


def _registry_port_forward(port):
    return _RegistryTunnel(port)


def _build_push(project, dockerfile, proxy_port):
    image = os.getenv("ADAPTDL_SUBMIT_REPO")
    external = image is not None
    if not external:
        image = "localhost:{}/adaptdl-submit".format(proxy_port)
    build = ["docker", "build", "-t", image, project]
    if dockerfile:
        build += ["-f", dockerfile]
    subprocess.check_call(build)
    forward = None if external else _registry_port_forward(proxy_port)
    try:
        subprocess.check_call(["docker", "push", image])
    finally:
        if forward is not None:
            forward.terminate()
    digests = json.loads(subprocess.check_output(
        ["docker", "image", "inspect", image,
         "--format={{json .RepoDigests}}"]))
    for digest in digests:
        if digest.startswith(image):
            remote = image if external else "localhost:32000/adaptdl-submit"
            return remote + "@" + digest.split("@")[-1]
    raise SystemExit("Error: no repo digest for {}".format(image))


def submit(args, remaining):
    image = _build_push(args.project, args.dockerfile, args.proxy_port)
    jobfile = args.jobfile or os.path.join(args.project, "adaptdljob.yaml")
    resource = _load_yaml(jobfile)
    if args.tensorboard is not None:
        _kubectl("get", "pvc", manifests.TENSORBOARD_PREFIX
                 + args.tensorboard)
    job, pvc_name = manifests.prepare_job(
        resource, image, remaining, name=args.name,
        pull_secret=os.getenv("ADAPTDL_SUBMIT_REPO_CREDS"),
        tensorboard=args.tensorboard)
    if args.pod_per_node:
        manifests.use_node_pods(job)
    created = _create(job)
    classes = json.loads(_kubectl("get", "storageclass", "-o",
                                  "json"))["items"]
    storage_class = manifests.choose_storageclass(
        classes, args.checkpoint_storage_class)
    _create(manifests.pvc_manifest(pvc_name, storage_class,
                                   args.checkpoint_storage_size,
                                   created["metadata"]))
    print("job {} submitted".format(created["metadata"]["name"]))


def cp(args, remaining):
    job_name, _, remote_path = args.source.partition(":")
    if not remote_path.startswith("/"):
        raise SystemExit("absolute path is required for the source path")
    job = json.loads(_kubectl("get", "adaptdljobs", job_name, "-o", "json"))
    for volume in job["spec"]["template"]["spec"].get("volumes", []):
        if volume["name"] == manifests.PVC_VOLUME:
            pvc_name = volume["persistentVolumeClaim"]["claimName"]
            break
    else:
        raise SystemExit("Error: job {} has no AdaptDL volume".format(
            job_name))
    pod = _create(manifests.copy_pod_manifest(pvc_name,
                                              str(uuid.uuid4())[:8]))
    name = pod["metadata"]["name"]
    try:
        _kubectl("wait", "--for=condition=Ready", "pod/" + name,
                 "--timeout=120s")
        _kubectl("cp", "{}:/adaptdl_pvc{}".format(name, remote_path),
                 args.destination, capture=False)
    finally:
        _kubectl("delete", "pod", name, "--wait=false")


def logs(args, remaining):
    while True:
        try:
            subprocess.check_call(
                ["kubectl", "logs", "-l",
                 "adaptdl/job={}".format(args.jobname)] + remaining)
            return
        except KeyboardInterrupt:
            return
        except subprocess.CalledProcessError:
            print("PRESS CTRL-C TO EXIT....")
            time.sleep(2)


def ls(args, remaining):
    items = json.loads(_kubectl("get", "adaptdljobs", "-o", "json"))["items"]
    if not items:
        print("No adaptdljobs")
        return
    rows = manifests.summarize_jobs(items, datetime.utcnow())
    fmt = "{:<40} {:<10} {:<20} {:<12} {:<9} {:<8}"
    print(fmt.format("Name", "Status", "Start(UTC)", "Runtime", "Replicas",
                     "Restarts"))
    for row in rows:
        print(fmt.format(row["name"], row["phase"],
                         row["start_time"].strftime("%Y-%m-%d %H:%M:%S"),
                         row["run_time"], str(row["replicas"]),
                         str(row["restarts"])))


def tensorboard(args, remaining):
    if args.tb_command == "create":
        classes = json.loads(_kubectl("get", "storageclass", "-o",
                                      "json"))["items"]
        sc = manifests.choose_storageclass(classes, args.storage_class)
        for obj in manifests.tensorboard_manifests(args.name, sc, args.size,
                                                   nodeport=args.nodeport):
            _create(obj)
    elif args.tb_command == "delete":
        full = manifests.TENSORBOARD_PREFIX + args.name
        for kind in ("deployment", "service", "pvc"):
            _kubectl("delete", kind, full, "--ignore-not-found")
    elif args.tb_command == "list":
        out = json.loads(_kubectl(
            "get", "deployment", "-l", "app=adaptdl-tensorboard", "-o",
            "json"))
        for item in out["items"]:
            print(item["metadata"]["labels"]["adaptdl/tensorboard"])
    elif args.tb_command == "proxy":
        full = manifests.TENSORBOARD_PREFIX + args.name
        print("TensorBoard at http://{}:{}".format(args.address, args.port))
        try:
            from adaptdl_b200.cli.proxy import service_proxy
            namespace = _kubectl(
                "config", "view", "--minify", "-o",
                "jsonpath={..namespace}").strip() or "default"
            with service_proxy(namespace, full + ":6006",
                               listen_host=args.address,
                               listen_port=args.port):
                threading.Event().wait()        # until interrupted
        except KeyboardInterrupt:
            pass
        except Exception:  # noqa: BLE001 - e.g. exec credential plugins
            _kubectl("port-forward", "--address", args.address,
                     "service/" + full, "{}:6006".format(args.port),
                     capture=False)


def build_parser():
    name = os.path.basename(sys.argv[0] or "") if sys.argv else ""
    # installed under both names (setup.py); "adaptdl" is the reference's
    parser = argparse.ArgumentParser(
        prog=name if name in ("adaptdl", "adaptdl-b200") else "adaptdl-b200")
    sub = parser.add_subparsers(dest="command", required=True)
    p = sub.add_parser("submit", help="build, push and submit a job")
    p.add_argument("project", help="directory with the job's Dockerfile")
    p.add_argument("-f", "--jobfile")
    p.add_argument("-d", "--dockerfile")
    p.add_argument("-n", "--name")
    p.add_argument("--proxy-port", type=int, default=59283)
    p.add_argument("--tensorboard")
    p.add_argument("--checkpoint-storage-class")
    p.add_argument("--checkpoint-storage-size", default="1Gi")
    p.add_argument("--pod-per-node", action="store_true",
                   help="one pod per node hosting all of the job's replicas "
                        "there (fused peer-memory reducer); wraps python "
                        "commands in the replica launcher")
    p.set_defaults(handler=submit)
    p = sub.add_parser("logs", help="stream a job's logs")
    p.add_argument("jobname")
    p.set_defaults(handler=logs)
    p = sub.add_parser("ls", help="list jobs")
    p.set_defaults(handler=ls)
    p = sub.add_parser("cp", help="copy files out of a job's volume")
    p.add_argument("source", help="<jobname>:/absolute/path")
    p.add_argument("destination")
    p.set_defaults(handler=cp)
    p = sub.add_parser("tensorboard", help="manage TensorBoard instances")
    tb = p.add_subparsers(dest="tb_command", required=True)
    c = tb.add_parser("create")
    c.add_argument("name")
    c.add_argument("--storage-class", "--storageclass",
                   dest="storage_class")      # second spelling: reference's
    c.add_argument("--size", default="1Gi")
    c.add_argument("--nodeport", action="store_true",
                   help="expose TensorBoard on a node port")
    for verb in ("delete", "proxy"):
        c = tb.add_parser(verb)
        c.add_argument("name")
        if verb == "proxy":
            c.add_argument("--address", default="127.0.0.1",
                           help="local address to bind")
            c.add_argument("-p", "--port", type=int, default=6006)
    tb.add_parser("list")
    p.set_defaults(handler=tensorboard)
    return parser


def main(argv=None):
    parser = build_parser()
    args, remaining = parser.parse_known_args(argv)
    if remaining and remaining[0] == "--":
        remaining = remaining[1:]
    args.handler(args, remaining)


if __name__ == "__main__":
    sys.exit(main())
