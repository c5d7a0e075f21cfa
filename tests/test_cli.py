"""CLI manifest construction (pure functions)."""
import json
import os
import sys
from datetime import datetime

import pytest

from adaptdl_b200.cli import manifests
from adaptdl_b200.cli.main import build_parser

JOB = {"apiVersion": "adaptdl.petuum.com/v1", "kind": "AdaptDLJob",
       "metadata": {"name": "cifar"},
       "spec": {"maxReplicas": 8, "template": {"spec": {"containers": [
           {"name": "main", "command": ["python3", "main.py"]}]}}}}


def test_prepare_job_injects_volumes_and_env():
    job, pvc = manifests.prepare_job(JOB, "repo/img@sha256:abc",
                                     ["--epochs", "3"], name="run",
                                     pull_secret="creds", tensorboard="tb1",
                                     pvc_name="adaptdl-pvc-x")
    assert pvc == "adaptdl-pvc-x"
    assert job["metadata"] == {"generateName": "run-"}
    spec = job["spec"]["template"]["spec"]
    main = spec["containers"][0]
    assert main["image"] == "repo/img@sha256:abc"
    assert main["args"] == ["--epochs", "3"]
    assert spec["imagePullSecrets"] == [{"name": "creds"}]
    env = {e["name"]: e["value"] for e in main["env"]}
    assert env == {"ADAPTDL_TENSORBOARD_LOGDIR": "/adaptdl/tensorboard",
                   "ADAPTDL_CHECKPOINT_PATH": "/adaptdl/checkpoint",
                   "ADAPTDL_SHARE_PATH": "/adaptdl/share"}
    mounts = {m["mountPath"]: m for m in main["volumeMounts"]}
    assert mounts["/adaptdl/checkpoint"]["subPath"] == "adaptdl/checkpoint"
    assert mounts["/adaptdl/share"]["subPath"] == "adaptdl/share"
    claims = {v["name"]: v["persistentVolumeClaim"]["claimName"]
              for v in spec["volumes"]}
    assert claims == {"adaptdl-tensorboard": "adaptdl-tensorboard-tb1",
                      "adaptdl-pvc": "adaptdl-pvc-x"}
    assert "volumes" not in JOB["spec"]["template"]["spec"]   # not mutated


def test_storageclass_choice_and_manifests():
    classes = [
        {"metadata": {"name": "gp2", "annotations": {
            "storageclass.kubernetes.io/is-default-class": "true"}},
         "provisioner": "kubernetes.io/aws-ebs"},
        {"metadata": {"name": "efs"}, "provisioner": "efs.csi.aws.com"}]
    assert manifests.choose_storageclass(classes) == "efs"
    assert manifests.choose_storageclass(classes[:1]) == "gp2"
    assert manifests.choose_storageclass(classes, "gp2") == "gp2"
    with pytest.raises(SystemExit):
        manifests.choose_storageclass(classes, "nope")
    pvc = manifests.pvc_manifest("p", "efs", "5Gi",
                                 {"name": "job", "uid": "u"})
    assert pvc["spec"]["accessModes"] == ["ReadWriteMany"]
    assert pvc["metadata"]["ownerReferences"][0]["kind"] == "AdaptDLJob"
    pod = manifests.copy_pod_manifest("p", "1234")
    assert pod["spec"]["containers"][0]["volumeMounts"][0]["mountPath"] \
        == "/adaptdl_pvc"
    objs = manifests.tensorboard_manifests("exp", "efs")
    assert [o["kind"] for o in objs] == ["PersistentVolumeClaim",
                                         "Deployment", "Service"]
    assert objs[1]["metadata"]["name"] == "adaptdl-tensorboard-exp"


def test_ls_summary_and_parser():
    items = [{"metadata": {"name": "a",
                           "creationTimestamp": "2026-01-01T00:00:00Z"},
              "status": {"phase": "Running", "replicas": 4, "group": 2}},
             {"metadata": {"name": "b",
                           "creationTimestamp": "2026-01-01T00:00:00Z"},
              "status": {"phase": "Succeeded",
                         "completionTimestamp":
                             "2026-01-01T01:30:00.000000+00:00"}}]
    rows = manifests.summarize_jobs(items, datetime(2026, 1, 1, 0, 10))
    assert rows[0]["run_time"] == "0:10:00" and rows[0]["restarts"] == 2
    assert rows[1]["run_time"] == "1:30:00" and rows[1]["replicas"] == "N/A"
    args, rest = build_parser().parse_known_args(
        ["submit", "proj", "--tensorboard", "tb", "--", "--lr", "0.1"])
    assert args.project == "proj" and rest[-2:] == ["--lr", "0.1"]
    args, _ = build_parser().parse_known_args(["tensorboard", "proxy", "x"])
    assert args.tb_command == "proxy" and args.port == 6006


def test_workload_specs_validate():
    """Every workload of the suite (tests/workloads) becomes a job the CLI
    can prepare and the scheduler's validator accepts."""
    import asyncio
    import importlib.util
    import os
    from adaptdl_b200.cli import manifests
    from adaptdl_b200.sched.kube import InMemoryCluster
    from adaptdl_b200.sched.validator import Validator
    path = os.path.join(os.path.dirname(__file__), "workloads",
                        "workloads.py")
    spec = importlib.util.spec_from_file_location("workloads", path)
    workloads = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(workloads)
    assert len(workloads.WORKLOADS) >= 14
    validator = Validator(InMemoryCluster())
    for name in workloads.WORKLOADS:
        resource = workloads.manifest(name)
        command = resource["spec"]["template"]["spec"]["containers"][0][
            "command"]
        script = next(part for part in command if part.endswith(".py"))
        assert ("adaptdl_b200.launch" in command) == bool(
            resource["spec"].get("podPerNode"))
        assert os.path.exists(os.path.join(
            workloads.ROOT, os.path.relpath(script, workloads.IMAGE_ROOT))), script
        job, pvc = manifests.prepare_job(resource, "registry/img@sha256:0",
                                         [], name=name)
        review = {"operation": "CREATE", "namespace": "default",
                  "uid": "u", "object": job}
        response = asyncio.run(validator._validate_create(review))
        assert response["allowed"], (name, response)
        assert os.path.exists(workloads.local_command(name)[0])


def test_jobfile_from_stdin(monkeypatch):
    import io
    from adaptdl_b200.cli import main as cli
    monkeypatch.setattr("sys.stdin", io.StringIO("kind: AdaptDLJob\nspec: {}\n"))
    assert cli._load_yaml("-") == {"kind": "AdaptDLJob", "spec": {}}


def test_pod_per_node_flag_wraps_python_commands():
    from adaptdl_b200.cli import manifests
    from adaptdl_b200.cli.main import build_parser
    args, rest = build_parser().parse_known_args(
        ["submit", ".", "--pod-per-node"])
    assert args.pod_per_node
    job = {"spec": {"template": {"spec": {"containers": [
        {"name": "main", "command": ["python3", "train.py", "--x"]},
        {"name": "side", "command": ["/bin/sidecar"]},
        {"name": "img-entrypoint"}]}}}}
    manifests.use_node_pods(job)
    manifests.use_node_pods(job)                       # idempotent
    containers = job["spec"]["template"]["spec"]["containers"]
    assert job["spec"]["podPerNode"] is True
    assert containers[0]["command"] == ["python3", "-m",
                                        "adaptdl_b200.launch", "train.py",
                                        "--x"]
    assert containers[1]["command"] == ["/bin/sidecar"]
    assert "command" not in containers[2]


def test_service_proxy_relays_through_the_api_server():
    """``service_proxy``: requests to the local port arrive at
    ``<api-server>/api/v1/namespaces/<ns>/services/<svc>/proxy<path>`` with
    the kubeconfig's bearer token, bodies and status codes pass through
    (reference: cli/adaptdl_cli/proxy.py:30-70)."""
    import json
    import threading
    import urllib.request
    from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
    from adaptdl_b200.cli.proxy import load_kube_access, service_proxy
    seen = []

    class ApiServer(BaseHTTPRequestHandler):
        def log_message(self, *args):
            pass

        def _any(self):
            n = int(self.headers.get("Content-Length") or 0)
            body = self.rfile.read(n) if n else b""
            seen.append((self.command, self.path,
                         self.headers.get("Authorization"), body))
            payload = json.dumps({"path": self.path,
                                  "echo": body.decode()}).encode()
            self.send_response(201 if self.command == "PUT" else 200)
            self.send_header("Content-Type", "application/json")
            self.send_header("Docker-Distribution-Api-Version", "registry/2.0")
            self.send_header("Content-Length", str(len(payload)))
            self.end_headers()
            self.wfile.write(payload)
        do_GET = do_PUT = _any

    api = ThreadingHTTPServer(("127.0.0.1", 0), ApiServer)
    threading.Thread(target=api.serve_forever, daemon=True).start()
    try:
        access = load_kube_access({
            "clusters": [{"cluster": {
                "server": "http://127.0.0.1:{}".format(
                    api.server_address[1])}}],
            "users": [{"user": {"token": "s3cret"}}]})
        with service_proxy("ns1", "adaptdl-registry:registry",
                           access=access) as addr:
            with urllib.request.urlopen(
                    "http://{}/v2/_catalog?n=5".format(addr)) as resp:
                assert resp.status == 200
                assert resp.headers["Docker-Distribution-Api-Version"]
                got = json.loads(resp.read())
            assert got["path"] == ("/api/v1/namespaces/ns1/services/"
                                   "adaptdl-registry:registry/proxy"
                                   "/v2/_catalog?n=5")
            req = urllib.request.Request(
                "http://{}/v2/blob".format(addr), data=b"layer-bytes",
                method="PUT")
            with urllib.request.urlopen(req) as resp:
                assert resp.status == 201
                assert json.loads(resp.read())["echo"] == "layer-bytes"
        assert [s[0] for s in seen] == ["GET", "PUT"]
        assert all(s[2] == "Bearer s3cret" for s in seen)
    finally:
        api.shutdown()
        api.server_close()


def test_kube_access_materialises_inline_credentials():
    import base64
    import os
    from adaptdl_b200.cli.proxy import load_kube_access
    blob = base64.b64encode(b"PEM").decode()
    access = load_kube_access({
        "clusters": [{"cluster": {"server": "https://k:6443/",
                                  "certificate-authority-data": blob}}],
        "users": [{"user": {"client-certificate-data": blob,
                            "client-key-data": blob}}]})
    try:
        assert access.server == "https://k:6443"
        assert open(access.verify, "rb").read() == b"PEM"
        assert all(open(p, "rb").read() == b"PEM" for p in access.cert)
        paths = [access.verify] + list(access.cert)
    finally:
        access.close()
    assert not any(os.path.exists(p) for p in paths)
    insecure = load_kube_access({"clusters": [{"cluster": {
        "server": "https://k", "insecure-skip-tls-verify": True}}]})
    assert insecure.verify is False and insecure.cert is None


def test_parser_accepts_the_reference_command_lines():
    """Every flag of the reference's ``adaptdl`` command (cli/bin/adaptdl:
    400-480, cli/adaptdl_cli/tensorboard.py:180-211) parses here, with the
    reference's spellings."""
    from adaptdl_b200.cli.main import build_parser
    parser = build_parser()
    a = parser.parse_args(["submit", "proj", "-f", "job.yaml", "-d",
                           "Dockerfile", "-n", "name", "--tensorboard", "tb",
                           "--checkpoint-storage-size", "2Gi",
                           "--checkpoint-storage-class", "efs",
                           "--proxy-port", "1234"])
    assert (a.project, a.jobfile, a.dockerfile, a.name, a.tensorboard) == \
        ("proj", "job.yaml", "Dockerfile", "name", "tb")
    assert a.checkpoint_storage_size == "2Gi" and a.proxy_port == 1234
    assert parser.parse_args(["logs", "job"]).jobname == "job"
    assert parser.parse_args(["ls"]).command == "ls"
    a = parser.parse_args(["cp", "job:/adaptdl/checkpoint/x", "out"])
    assert a.source == "job:/adaptdl/checkpoint/x" and a.destination == "out"
    a = parser.parse_args(["tensorboard", "create", "tb", "--nodeport",
                           "--storageclass", "fast", "--size", "5Gi"])
    assert a.storage_class == "fast" and a.nodeport and a.size == "5Gi"
    a = parser.parse_args(["tensorboard", "proxy", "tb", "--address",
                           "0.0.0.0", "-p", "7007"])
    assert a.address == "0.0.0.0" and a.port == 7007
    for verb in ("delete", "list"):
        parser.parse_args(["tensorboard", verb] + (["tb"] if verb == "delete"
                                                   else []))


def _run_cli(tmp_path, *argv, env_extra=None):
    """``python -m adaptdl_b200.cli.main ...`` with stand-ins for kubectl
    and docker in front of PATH (tests/fixtures/fake_bin); returns the
    process plus the recorded kubectl / docker invocations."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    klog, dlog = tmp_path / "kubectl.jsonl", tmp_path / "docker.jsonl"
    env = dict(os.environ, PYTHONPATH=root,
               PATH=os.path.join(here, "fixtures", "fake_bin") + os.pathsep
               + os.environ["PATH"],
               FAKE_KUBECTL_LOG=str(klog), FAKE_DOCKER_LOG=str(dlog),
               KUBECONFIG=str(tmp_path / "no-kubeconfig"))
    env.update(env_extra or {})
    proc = subprocess.run(
        [sys.executable, "-m", "adaptdl_b200.cli.main"] + list(argv),
        env=env, cwd=str(tmp_path), stdout=subprocess.PIPE,
        stderr=subprocess.STDOUT, text=True, timeout=120)

    def rows(path):
        if not path.exists():
            return []
        return [json.loads(line) for line in path.read_text().splitlines()]
    return proc, rows(klog), rows(dlog)


def test_submit_end_to_end_against_stand_in_kubectl_and_docker(tmp_path):
    """``adaptdl submit``: image built and pushed, digest-pinned image and
    the checkpoint volume in the job that is created, PVC owned by the job
    (capabilities of cli/bin/adaptdl:183-316)."""
    project = tmp_path / "proj"
    project.mkdir()
    (project / "Dockerfile").write_text("FROM scratch\n")
    (project / "adaptdljob.yaml").write_text(
        "apiVersion: adaptdl.petuum.com/v1\nkind: AdaptDLJob\n"
        "metadata: {generateName: demo-}\n"
        "spec:\n  template:\n    spec:\n      containers:\n"
        "      - {name: main, command: [python3, train.py]}\n")
    proc, kube, docker = _run_cli(
        tmp_path, "submit", str(project), "--checkpoint-storage-size", "2Gi",
        env_extra={"ADAPTDL_SUBMIT_REPO": "registry.example/team/img"})
    assert proc.returncode == 0, proc.stdout
    assert "submitted" in proc.stdout
    assert [d[:2] for d in docker] == [["build", "-t"], ["push",
                                                         "registry.example"
                                                         "/team/img"],
                                       ["image", "inspect"]]
    created = [json.loads(k["stdin"]) for k in kube
               if k["argv"][:1] == ["create"]]
    job, pvc = created
    assert job["kind"] == "AdaptDLJob"
    container = job["spec"]["template"]["spec"]["containers"][0]
    assert container["image"] == "registry.example/team/img@sha256:" \
        + "0" * 64
    env = {e["name"]: e.get("value") for e in container["env"]}
    assert env["ADAPTDL_CHECKPOINT_PATH"] == "/adaptdl/checkpoint"
    assert pvc["kind"] == "PersistentVolumeClaim"
    assert pvc["spec"]["storageClassName"] == "efs"      # cluster default
    assert pvc["spec"]["resources"]["requests"]["storage"] == "2Gi"
    assert pvc["metadata"]["ownerReferences"][0]["kind"] == "AdaptDLJob"
    claim = [v for v in job["spec"]["template"]["spec"]["volumes"]
             if v["name"] == "adaptdl-pvc"][0]
    assert claim["persistentVolumeClaim"]["claimName"] == \
        pvc["metadata"]["name"]


def test_ls_logs_cp_and_tensorboard_commands_end_to_end(tmp_path):
    proc, kube, _ = _run_cli(tmp_path, "ls")
    assert proc.returncode == 0, proc.stdout
    assert "job-abc" in proc.stdout and "Running" in proc.stdout

    proc, kube, _ = _run_cli(tmp_path, "logs", "job-abc", "-f")
    assert proc.returncode == 0, proc.stdout
    assert "line from replica 0" in proc.stdout
    assert kube[-1]["argv"] == ["logs", "-l", "adaptdl/job=job-abc", "-f"]

    proc, kube, _ = _run_cli(tmp_path, "cp", "job-abc:/adaptdl/checkpoint/x",
                             "out")
    assert proc.returncode == 0, proc.stdout
    verbs = [k["argv"][0] for k in kube if k["argv"][0] in
             ("get", "create", "wait", "cp", "delete")][-5:]
    assert verbs == ["get", "create", "wait", "cp", "delete"], kube
    copy_pod = json.loads([k for k in kube
                           if k["argv"][0] == "create"][-1]["stdin"])
    volume = copy_pod["spec"]["volumes"][0]["persistentVolumeClaim"]
    assert volume["claimName"] == "adaptdl-pvc-1234"
    cp_call = [k["argv"] for k in kube if k["argv"][0] == "cp"][-1]
    assert cp_call[1].endswith(":/adaptdl_pvc/adaptdl/checkpoint/x")
    assert cp_call[2] == "out"
    proc, _, _ = _run_cli(tmp_path, "cp", "job-abc:relative/path", "out")
    assert proc.returncode != 0 and "absolute path" in proc.stdout

    proc, kube, _ = _run_cli(tmp_path, "tensorboard", "create", "tb1",
                             "--storageclass", "gp2", "--size", "3Gi",
                             "--nodeport")
    assert proc.returncode == 0, proc.stdout
    kinds = [json.loads(k["stdin"])["kind"] for k in kube
             if k["argv"][0] == "create"][-3:]
    assert sorted(kinds) == ["Deployment", "PersistentVolumeClaim",
                             "Service"]
    proc, _, _ = _run_cli(tmp_path, "tensorboard", "list")
    assert proc.returncode == 0 and "tb1" in proc.stdout
    proc, kube, _ = _run_cli(tmp_path, "tensorboard", "delete", "tb1")
    assert proc.returncode == 0, proc.stdout
    deleted = [k["argv"][1:3] for k in kube if k["argv"][0] == "delete"][-3:]
    assert deleted == [["deployment", "adaptdl-tensorboard-tb1"],
                       ["service", "adaptdl-tensorboard-tb1"],
                       ["pvc", "adaptdl-tensorboard-tb1"]]


def test_submit_through_the_in_cluster_registry(tmp_path):
    """Without ``ADAPTDL_SUBMIT_REPO`` the image goes to the chart's
    insecure registry through a local tunnel (API-server proxy, or ``kubectl
    port-forward`` when the kubeconfig cannot be used -- the case here) and
    the job pulls it from the node-local address."""
    project = tmp_path / "proj"
    project.mkdir()
    (project / "Dockerfile").write_text("FROM scratch\n")
    (project / "job.yaml").write_text(
        "apiVersion: adaptdl.petuum.com/v1\nkind: AdaptDLJob\n"
        "metadata: {generateName: demo-}\n"
        "spec:\n  template:\n    spec:\n      containers:\n"
        "      - {name: main, command: [python3, train.py]}\n")
    proc, kube, docker = _run_cli(
        tmp_path, "submit", str(project), "-f", str(project / "job.yaml"),
        "--proxy-port", "59999", "--pod-per-node", "--", "--epochs", "3")
    assert proc.returncode == 0, proc.stdout
    assert docker[0][:3] == ["build", "-t", "localhost:59999/adaptdl-submit"]
    assert docker[1] == ["push", "localhost:59999/adaptdl-submit"]
    assert any(k["argv"][:2] == ["port-forward", "service/adaptdl-registry"]
               for k in kube), kube
    job = json.loads([k for k in kube
                      if k["argv"][:1] == ["create"]][0]["stdin"])
    container = job["spec"]["template"]["spec"]["containers"][0]
    assert container["image"].startswith(
        "localhost:32000/adaptdl-submit@sha256:")
    assert job["spec"]["podPerNode"] is True
    # python commands are wrapped in the replica launcher, user arguments
    # are appended
    assert container["command"][:4] == ["python3", "-m",
                                        "adaptdl_b200.launch", "train.py"]
    assert (container.get("args") or container["command"])[-2:] == \
        ["--epochs", "3"]


def test_workload_suite_submits_through_the_cli_and_soaks(tmp_path):
    """tests/workloads/workloads.py (the reference's tests/*-workload/*.sh
    and testworkload.sh as one table-driven script): ``submit`` pipes the
    generated AdaptDLJob into ``adaptdl submit -f -``, ``soak`` tops the
    cluster up to N active jobs -- against the stand-in kubectl / docker."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    klog, dlog = tmp_path / "kubectl.jsonl", tmp_path / "docker.jsonl"
    env = dict(os.environ, PYTHONPATH=root,
               PATH=os.path.join(here, "fixtures", "fake_bin") + os.pathsep
               + os.environ["PATH"],
               FAKE_KUBECTL_LOG=str(klog), FAKE_DOCKER_LOG=str(dlog),
               ADAPTDL_SUBMIT_REPO="registry.example/team/img")
    script = os.path.join(here, "workloads", "workloads.py")

    def run(*argv):
        return subprocess.run([sys.executable, script] + list(argv), env=env,
                              cwd=str(tmp_path), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True,
                              timeout=300)
    proc = run("submit", "resnet18-cifar10-short")
    assert proc.returncode == 0 and "submitted" in proc.stdout, proc.stdout
    created = [json.loads(line) for line in klog.read_text().splitlines()]
    job = json.loads([k for k in created
                      if k["argv"][:1] == ["create"]][0]["stdin"])
    command = job["spec"]["template"]["spec"]["containers"][0]["command"]
    assert any(part.endswith("pytorch-cifar/main.py") for part in command)
    # the stand-in cluster always shows one active job: two wanted -> one
    # more is submitted per round
    proc = run("soak", "--suite", "short", "--jobs", "2", "--rounds", "1",
               "--period", "0")
    assert proc.returncode == 0, proc.stdout
    assert proc.stdout.count("submitting") == 1, proc.stdout
