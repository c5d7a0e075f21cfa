"""Packaging. The CUDA extension is NOT a setuptools ext_module: it is built
in-tree by ``adaptdl_b200._native.build()`` (plain nvcc, sm_100a) so the same
``.so`` ships with a repo snapshot; ``python setup.py build_native`` runs it."""
import os

import shutil

import setuptools
from setuptools import Command
from setuptools.command.build_py import build_py


class BuildNative(Command):
    description = "compile csrc/ for sm_100a into adaptdl_b200/_native"
    user_options = []

    def initialize_options(self):
        pass

    def finalize_options(self):
        pass

    def run(self):
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from adaptdl_b200 import _native
        from adaptdl_b200._native import host
        print(host.build(force=True))        # g++ only (scheduler pods)
        print(_native.build(force=True))     # nvcc, sm_100a


class BuildPyWithCsrc(build_py):
    """Installed packages carry the native sources (``csrc/``: CUDA, and
    ``csrc/host``: plain C++) inside ``adaptdl_b200/_native/csrc`` so
    ``adaptdl_b200._native.build()`` / ``_native.host.build()`` can compile
    them on the target machine."""

    def run(self):
        super().run()
        here = os.path.dirname(os.path.abspath(__file__))
        target = os.path.join(self.build_lib, "adaptdl_b200", "_native",
                              "csrc")
        os.makedirs(target, exist_ok=True)
        for sub in ("", "host"):
            os.makedirs(os.path.join(target, sub), exist_ok=True)
            for name in sorted(os.listdir(os.path.join(here, "csrc", sub))):
                if name.endswith((".cu", ".cuh", ".cpp", ".h")):
                    shutil.copy2(os.path.join(here, "csrc", sub, name),
                                 os.path.join(target, sub))


setuptools.setup(
    name="adaptdl-b200",
    version=os.getenv("ADAPTDL_VERSION", "0.1.0"),
    description="Blackwell-native elastic data-parallel training engine with "
                "adaptive batch size / learning rate and goodput-driven "
                "scheduling",
    # adaptdl / adaptdl_sched / adaptdl_ray / adaptdl_cli: three-line alias
    # packages that resolve the reference's import names to this framework
    # (adaptdl_b200/compat.py), so scripts and manifests written for
    # petuum/adaptdl keep working after `pip install`
    packages=setuptools.find_packages(include=[
        "adaptdl_b200", "adaptdl_b200.*", "adaptdl", "adaptdl_sched",
        "adaptdl_ray", "adaptdl_cli"]),
    package_data={"adaptdl_b200._native": ["*.so"]},
    python_requires=">=3.9",
    install_requires=["numpy", "scipy", "torch>=2.1", "requests"],
    extras_require={
        "sched": ["aiohttp", "prometheus_client", "kubernetes_asyncio",
                  "pyyaml"],
        "ray": ["ray[tune]"],
    },
    entry_points={"console_scripts": [
        "adaptdl-b200=adaptdl_b200.cli.main:main",
        # the reference's command names (cli/bin/adaptdl, ray/setup.py:49)
        "adaptdl=adaptdl_b200.cli.main:main",
        "adaptdl_on_ray_aws=adaptdl_b200.ray.aws.launch_job:main",
        "adaptdl-b200-launch=adaptdl_b200.launch:main",
        "adaptdl-b200-local=adaptdl_b200.sched.local:main",
        "adaptdl-b200-local-cluster=adaptdl_b200.sched.local_cluster:main",
        "adaptdl-b200-on-ray-aws=adaptdl_b200.ray.aws.launch_job:main",
    ]},
    cmdclass={"build_native": BuildNative, "build_py": BuildPyWithCsrc},
)
