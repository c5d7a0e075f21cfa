"""Admission webhook for AdaptDLJob objects.

Kubernetes sends every CREATE / UPDATE of an AdaptDLJob to ``POST /validate``
as an ``AdmissionReview`` and acts on the verdict:

* CREATE -- the job's pod template must be something the API server would
  accept (checked with a dry-run pod creation, so image pull secrets, volume
  references, resource syntax, ... fail at submission time instead of when
  the scheduler first starts the job), and ``maxReplicas`` must not be
  smaller than ``minReplicas``;
* UPDATE -- a job's ``spec`` is immutable; only ``status`` (written by the
  scheduler) and metadata may change.

The rules are plain functions (:func:`verdict_for_create`,
:func:`verdict_for_update`) wrapped by a small aiohttp application
(capabilities of the reference's ``sched/adaptdl_sched/validator.py:30-134``).
"""

import argparse
import json
import logging
import ssl
from http import HTTPStatus

from aiohttp import web

from adaptdl_b200.sched.kube import ApiError

LOG = logging.getLogger(__name__)

ALLOW = {"allowed": True}


def refuse(reason, message):
    """An admission response that rejects the request (HTTP 422)."""
    status = {"code": HTTPStatus.UNPROCESSABLE_ENTITY.value,
              "reason": reason, "message": message}
    return {"allowed": False, "status": status}


def pod_of_template(job):
    """A stand-alone Pod object made from the job's pod template."""
    template = job["spec"].get("template") or {}
    metadata = dict(template.get("metadata") or {})
    metadata["name"] = "spec.template"       # shows up in API error messages
    return {"apiVersion": "v1", "kind": "Pod", "metadata": metadata,
            "spec": template.get("spec") or {}}


async def verdict_for_create(cluster, namespace, job):
    pod = pod_of_template(job)
    if isinstance(cluster, _TemplateDryRun):
        pod["_template"] = job["spec"].get("template")
    try:
        await cluster.create_pod(namespace, pod, dry_run=True)
    except ApiError as exc:
        return refuse("Invalid", exc.message
                      if isinstance(cluster, _TemplateDryRun) else str(exc))
    spec = job["spec"]
    lowest = spec.get("minReplicas", 0)
    highest = spec.get("maxReplicas")
    if highest is not None and highest < lowest:
        return refuse("Invalid", "spec.maxReplicas ({}) is smaller than "
                                 "spec.minReplicas ({})".format(highest,
                                                                lowest))
    return dict(ALLOW)


def verdict_for_update(old_job, new_job):
    if old_job["spec"] == new_job["spec"]:
        return dict(ALLOW)
    return refuse("Forbidden", "the spec of an AdaptDLJob cannot be changed "
                               "after creation")


class _TemplateDryRun(object):
    """``create_pod``-shaped front of the Kubernetes core API for a
    :class:`Validator` that was built without a cluster backend: the job's
    pod template is submitted as a dry-run ``PodTemplate`` object (nothing is
    persisted), the API server's own message is what the user gets back."""

    def __init__(self):
        import kubernetes_asyncio as kubernetes
        self._kubernetes = kubernetes
        self.core_api = kubernetes.client.CoreV1Api()

    async def create_pod(self, namespace, pod, dry_run=True):
        body = {"metadata": {"name": pod["metadata"].get("name")},
                "template": pod.pop("_template")}
        try:
            await self.core_api.create_namespaced_pod_template(
                namespace, body, dry_run="All")
        except self._kubernetes.client.rest.ApiException as exc:
            message = str(exc)
            try:
                message = json.loads(exc.body)["message"]
            except (AttributeError, KeyError, TypeError, ValueError):
                pass
            raise ApiError(exc.status, message)


class Validator(object):
    """The webhook server. ``cluster`` is a :mod:`adaptdl_b200.sched.kube`
    backend (only ``create_pod(..., dry_run=True)`` is used); without one
    (``Validator()``, as the reference constructs it) the checks go to the
    in-cluster Kubernetes API directly (needs ``kubernetes_asyncio``)."""

    def __init__(self, cluster=None):
        if cluster is None:
            cluster = _TemplateDryRun()
            self._core_api = cluster.core_api
        self._cluster = cluster
        self._app = web.Application()
        self._app.router.add_get("/healthz", self._healthz)
        self._app.router.add_post("/validate", self._validate)

    def get_app(self):
        return self._app

    def run(self, host, port, ssl_context=None):
        web.run_app(self._app, host=host, port=port, ssl_context=ssl_context)

    # kept as methods: tests and the CLI's manifest check call them directly
    async def _validate_create(self, review):
        return await verdict_for_create(self._cluster, review["namespace"],
                                        review["object"])

    @staticmethod
    def _validate_update(review):
        return verdict_for_update(review["oldObject"], review["object"])

    async def _healthz(self, request):
        return web.Response()

    async def _validate(self, request):
        review = (await request.json())["request"]
        kind = review["operation"]
        if kind == "CREATE":
            verdict = await self._validate_create(review)
        elif kind == "UPDATE":
            verdict = self._validate_update(review)
        else:                                  # DELETE / CONNECT: not ours
            verdict = dict(ALLOW)
        LOG.info("%s of %s/%s -> %s", kind, review.get("namespace"),
                 review.get("name", "<new>"), verdict)
        return web.json_response({
            "apiVersion": "admission.k8s.io/v1", "kind": "AdmissionReview",
            "response": dict(verdict, uid=review["uid"])})


def _load_cluster_credentials():
    """The webhook asks the API server for dry runs, so its client needs
    the pod's service-account credentials like the other scheduler
    containers (outside a cluster: whatever the client library defaults to,
    with a warning)."""
    import kubernetes_asyncio as kubernetes
    try:
        kubernetes.config.load_incluster_config()
    except Exception as exc:  # noqa: BLE001 - ConfigException outside a pod
        LOG.warning("no in-cluster credentials (%s); using the client "
                    "library's defaults", exc)


def main(argv=None):
    from adaptdl_b200.sched.kube import KubernetesCluster
    parser = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    parser.add_argument("--host", default="0.0.0.0")
    parser.add_argument("--port", type=int, default=8080)
    parser.add_argument("--tls-crt", help="server certificate (PEM)")
    parser.add_argument("--tls-key", help="private key (PEM)")
    args = parser.parse_args(argv)
    tls = None
    if args.tls_crt and args.tls_key:
        tls = ssl.SSLContext(ssl.PROTOCOL_TLS_SERVER)
        tls.load_cert_chain(args.tls_crt, args.tls_key)
    _load_cluster_credentials()
    Validator(KubernetesCluster()).run(args.host, args.port, tls)


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    main()
