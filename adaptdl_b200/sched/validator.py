"""Admission webhook for AdaptDLJob objects.

``POST /validate`` receives an ``AdmissionReview``: on CREATE the pod
template is dry-run-created against the API server and
``maxReplicas >= minReplicas`` is enforced; on UPDATE any change of
``spec`` is refused (jobs are immutable; the scheduler only writes
``status``). Parity: reference ``sched/adaptdl_sched/validator.py:30-134``.
"""

import argparse
import logging
import ssl
import sys
from http import HTTPStatus

from aiohttp import web

from adaptdl_b200.sched.kube import ApiError

LOG = logging.getLogger(__name__)


def _deny(reason, message):
    return {"allowed": False,
            "status": {"code": int(HTTPStatus.UNPROCESSABLE_ENTITY),
                       "reason": reason, "message": message}}


class Validator(object):

    def __init__(self, cluster):
        self._cluster = cluster
        self._app = web.Application()
        self._app.add_routes([
            web.get("/healthz", self._handle_healthz),
            web.post("/validate", self._handle_validate),
        ])

    def get_app(self):
        return self._app

    def run(self, host, port, ssl_context=None):
        web.run_app(self._app, host=host, port=port, ssl_context=ssl_context)

    async def _handle_healthz(self, request):
        return web.Response()

    async def _handle_validate(self, request):
        review = (await request.json())["request"]
        operation = review["operation"]
        if operation == "CREATE":
            response = await self._validate_create(review)
        elif operation == "UPDATE":
            response = self._validate_update(review)
        else:
            response = {"allowed": True}
        LOG.info("%s %s/%s: %s", operation, review.get("namespace"),
                 review.get("name", "<none>"), response)
        response["uid"] = review["uid"]
        return web.json_response({"apiVersion": "admission.k8s.io/v1",
                                  "kind": "AdmissionReview",
                                  "response": response})

    async def _validate_create(self, review):
        job = review["object"]
        template = job["spec"].get("template") or {}
        pod = {"apiVersion": "v1", "kind": "Pod",
               "metadata": dict(template.get("metadata") or {},
                                name="spec.template"),
               "spec": template.get("spec") or {}}
        try:
            await self._cluster.create_pod(review["namespace"], pod,
                                           dry_run=True)
        except ApiError as exc:
            return _deny("Invalid", str(exc))
        if job["spec"].get("maxReplicas", sys.maxsize) < \
                job["spec"].get("minReplicas", 0):
            return _deny("Invalid", "spec.maxReplicas must be greater than "
                                    "or equal to spec.minReplicas")
        return {"allowed": True}

    @staticmethod
    def _validate_update(review):
        if review["object"]["spec"] != review["oldObject"]["spec"]:
            return _deny("Forbidden", "updates to job spec are forbidden")
        return {"allowed": True}


def main(argv=None):
    from adaptdl_b200.sched.kube import KubernetesCluster
    parser = argparse.ArgumentParser()
    parser.add_argument("--host", default="0.0.0.0")
    parser.add_argument("--port", type=int, default=8080)
    parser.add_argument("--tls-crt")
    parser.add_argument("--tls-key")
    args = parser.parse_args(argv)
    context = None
    if args.tls_crt and args.tls_key:
        context = ssl.SSLContext(ssl.PROTOCOL_TLS_SERVER)
        context.load_cert_chain(args.tls_crt, args.tls_key)
    Validator(KubernetesCluster()).run(args.host, args.port, context)


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO)
    main()
