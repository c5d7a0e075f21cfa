"""Deployment artefacts stay consistent with the code that consumes them
(no helm / kubectl in the test environment: templates are checked as text
and plain manifests as YAML)."""
import glob
import os
import re

import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHART = os.path.join(ROOT, "deploy", "helm", "adaptdl-b200-sched")


def _read(*parts):
    with open(os.path.join(*parts)) as f:
        return f.read()


def _templates():
    return {os.path.basename(p): _read(p)
            for p in glob.glob(os.path.join(CHART, "templates", "*.yaml"))}


def test_every_values_reference_exists_in_values_yaml():
    values = yaml.safe_load(_read(CHART, "values.yaml"))
    missing = []
    for name, text in _templates().items():
        for path in set(re.findall(r"\.Values((?:\.[A-Za-z_]\w*)+)", text)):
            node = values
            for key in path.strip(".").split("."):
                if not isinstance(node, dict) or key not in node:
                    missing.append("{}: .Values{}".format(name, path))
                    break
                node = node[key]
    assert not missing, missing


def test_config_map_feeds_every_scheduler_setting():
    from adaptdl_b200.sched import config
    provided = set(re.findall(r'"(ADAPTDL_[A-Z_]+)"',
                              _templates()["config.yaml"]))
    provided |= set(re.findall(r"name: (ADAPTDL_[A-Z_]+)",
                               _templates()["sched.yaml"]))
    needed = {variable for variable, _, _ in config._SETTINGS.values()}
    assert needed <= provided, needed - provided
    # and the deployment runs exactly the roles the entry point knows
    main = _read(ROOT, "adaptdl_b200", "sched", "__main__.py")
    roles = re.search(r'list ((?:"\w+" ?)+)', _templates()["sched.yaml"])
    for role in re.findall(r'"(\w+)"', roles.group(1)):
        assert 'role == "{}"'.format(role) in main


def test_crd_matches_what_controller_and_cli_use():
    crd = yaml.safe_load(_templates()["crd.yaml"])
    from adaptdl_b200.sched import config
    assert crd["spec"]["group"] == config.GROUP
    assert crd["spec"]["names"]["plural"] == config.PLURAL
    version = crd["spec"]["versions"][0]
    assert version["name"] == config.VERSION
    assert "status" in version["subresources"]      # patch_job_status
    spec = version["schema"]["openAPIV3Schema"]["properties"]["spec"]
    assert set(spec["properties"]) == {"maxReplicas", "minReplicas",
                                       "preemptible", "template",
                                       "podPerNode"}
    assert spec["properties"]["maxReplicas"]["minimum"] == 1
    columns = [c["name"] for c in version["additionalPrinterColumns"]]
    assert columns == ["Ready", "Replicas", "Restarts", "Status", "Age"]


def test_plain_manifests_parse():
    for rel in ("deploy/cluster-autoscaler-values.yaml",
                "deploy/eks-cluster.yaml",
                "examples/ray/aws/cluster.yaml", "tutorial/mnist-job.yaml",
                ".github/workflows/test.yaml",
                ".github/workflows/docs.yaml",
                ".github/workflows/release.yaml"):
        docs = [d for d in yaml.safe_load_all(_read(ROOT, rel)) if d]
        assert docs, rel
    # the autoscaler discovers exactly the node groups eksctl tags
    values = yaml.safe_load(_read(ROOT, "deploy",
                                  "cluster-autoscaler-values.yaml"))
    eks = yaml.safe_load(_read(ROOT, "deploy", "eks-cluster.yaml"))
    assert values["autoDiscovery"]["clusterName"] == eks["metadata"]["name"]
    tagged = [group for group in eks["nodeGroups"]
              if all(tag in (group.get("tags") or {})
                     for tag in values["autoDiscovery"]["tags"])]
    assert [group["name"] for group in tagged] == ["gpu"]


@pytest.mark.parametrize("manifest", [("tutorial", "mnist-job.yaml"),
                                      ("examples", "BERT", "adaptdljob.yaml")])
def test_shipped_jobs_go_through_the_submit_path(manifest):
    from adaptdl_b200.cli import manifests
    resource = yaml.safe_load(_read(ROOT, *manifest))
    job, pvc = manifests.prepare_job(resource, "registry/img@sha256:0", [],
                                     name="tutorial")
    container = job["spec"]["template"]["spec"]["containers"][0]
    assert container["image"] == "registry/img@sha256:0"
    env = {e["name"]: e["value"] for e in container["env"]}
    assert env["ADAPTDL_CHECKPOINT_PATH"] == manifests.CHECKPOINT_MOUNT
    assert job["spec"]["maxReplicas"] >= job["spec"]["minReplicas"]
    assert pvc.startswith("adaptdl-pvc-")
    script = next(c for c in container["command"] if c.endswith(".py"))
    assert os.path.exists(os.path.join(ROOT, script))


def test_dockerfiles_copy_paths_that_exist():
    for rel in ("deploy/docker/Dockerfile.sched",
                "deploy/docker/Dockerfile.trainer", "examples/Dockerfile",
                "examples/BERT/Dockerfile", "tutorial/Dockerfile"):
        for line in _read(ROOT, rel).splitlines():
            if line.startswith("COPY "):
                for src in line.split()[1:-1]:
                    assert os.path.exists(os.path.join(ROOT, src)), \
                        (rel, src)


def test_distribution_ships_the_reference_import_names_and_commands():
    """``pip install`` must carry the alias packages (``import adaptdl`` ...)
    and the reference's command names, or a switched-over user finds neither
    after installing; images that build from a partial copy of the tree must
    copy them too."""
    import setuptools
    text = _read(ROOT, "setup.py")
    include = ["adaptdl_b200", "adaptdl_b200.*", "adaptdl", "adaptdl_sched",
               "adaptdl_ray", "adaptdl_cli"]
    for name in include:
        assert '"{}"'.format(name) in text, name
    found = set(setuptools.find_packages(ROOT, include=include))
    assert {"adaptdl", "adaptdl_sched", "adaptdl_ray", "adaptdl_cli",
            "adaptdl_b200.torch", "adaptdl_b200.sched.policy"} <= found
    for script in ('"adaptdl=adaptdl_b200.cli.main:main"',
                   '"adaptdl_on_ray_aws=adaptdl_b200.ray.aws.launch_job:main"'):
        assert script in text, script
    for rel in ("deploy/docker/Dockerfile.sched", "examples/Dockerfile",
                "examples/BERT/Dockerfile", "tutorial/Dockerfile"):
        copied = [line.split()[1] for line in _read(ROOT, rel).splitlines()
                  if line.startswith("COPY ")]
        if "." in copied:
            continue
        assert {"adaptdl", "adaptdl_sched", "adaptdl_ray",
                "adaptdl_cli"} <= set(copied), rel
