"""Local reverse proxy to a Kubernetes service through the API server.

``service_proxy(namespace, service)`` is the capability of the reference's
``cli/adaptdl_cli/proxy.py:30-70`` (a context manager that makes
``127.0.0.1:<port>`` talk to ``[https:]service[:port]`` in the cluster via the
API server's ``/proxy`` sub-resource, so that ``docker push`` reaches the
in-cluster registry and a browser reaches TensorBoard with nothing but the
user's kubeconfig). The reference embeds mitmproxy; this is ~100 lines of the
standard library plus ``requests``: a threaded HTTP server that replays each
request against ``<api-server>/api/v1/namespaces/<ns>/services/<svc>/proxy``
with the credentials of the current kubeconfig context, streaming bodies in
both directions.

The kubeconfig is read with ``kubectl config view --raw --minify -o json``
(no Kubernetes client library needed); client certificates / keys / CA given
inline (``*-data``) are written to private temporary files for the lifetime
of the proxy.
"""

import base64
import contextlib
import json
import os
import subprocess
import tempfile
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

__all__ = ["service_proxy", "load_kube_access", "KubeAccess"]

_HOP_BY_HOP = {"connection", "keep-alive", "proxy-authenticate",
               "proxy-authorization", "te", "trailers", "transfer-encoding",
               "upgrade", "host", "content-length"}


class KubeAccess(object):
    """How to reach the API server: base URL + ``requests`` keyword
    arguments (``headers``, ``cert``, ``verify``)."""

    def __init__(self, server, headers=None, cert=None, verify=True,
                 cleanup=()):
        self.server = server.rstrip("/")
        self.headers = dict(headers or {})
        self.cert = cert
        self.verify = verify
        self._cleanup = list(cleanup)

    def close(self):
        for path in self._cleanup:
            try:
                os.unlink(path)
            except OSError:
                pass
        self._cleanup = []


def _materialise(entry, name, tmp):
    """Path of ``entry[name]`` or of a temp file holding ``name-data``."""
    if entry.get(name):
        return entry[name]
    data = entry.get(name + "-data")
    if not data:
        return None
    handle = tempfile.NamedTemporaryFile("wb", delete=False,
                                         prefix="adaptdl-kube-")
    handle.write(base64.b64decode(data))
    handle.close()
    os.chmod(handle.name, 0o600)
    tmp.append(handle.name)
    return handle.name


def load_kube_access(config=None):
    """Credentials of the current kubectl context. ``config``: an already
    parsed ``kubectl config view --raw --minify -o json`` (tests)."""
    if config is None:
        config = json.loads(subprocess.check_output(
            ["kubectl", "config", "view", "--raw", "--minify", "-o",
             "json"]))
    cluster = config["clusters"][0]["cluster"]
    user = (config.get("users") or [{}])[0].get("user", {}) or {}
    tmp = []
    verify = True
    if cluster.get("insecure-skip-tls-verify"):
        verify = False
    else:
        ca = _materialise(cluster, "certificate-authority", tmp)
        if ca:
            verify = ca
    headers = {}
    if user.get("token"):
        headers["Authorization"] = "Bearer " + user["token"]
    cert = None
    client_cert = _materialise(user, "client-certificate", tmp)
    client_key = _materialise(user, "client-key", tmp)
    if client_cert and client_key:
        cert = (client_cert, client_key)
    return KubeAccess(cluster["server"], headers, cert, verify, tmp)


def _make_handler(access, prefix, verbose):
    import requests
    session = requests.Session()

    class Handler(BaseHTTPRequestHandler):
        protocol_version = "HTTP/1.1"

        def log_message(self, fmt, *args):
            if verbose:
                BaseHTTPRequestHandler.log_message(self, fmt, *args)

        def _relay(self):
            length = int(self.headers.get("Content-Length") or 0)
            body = self.rfile.read(length) if length else None
            headers = {k: v for k, v in self.headers.items()
                       if k.lower() not in _HOP_BY_HOP}
            # the service sees the address the CLIENT used (registries build
            # redirect / upload URLs from it)
            headers["Host"] = self.headers.get("Host", "")
            headers.update(access.headers)
            try:
                upstream = session.request(
                    self.command, access.server + prefix + self.path,
                    headers=headers, data=body, cert=access.cert,
                    verify=access.verify, stream=True,
                    allow_redirects=False, timeout=300)
            except requests.RequestException as exc:
                self.send_error(502, "upstream error: {}".format(exc))
                return
            payload = upstream.raw.read(decode_content=False)
            self.send_response(upstream.status_code)
            for key, value in upstream.headers.items():
                if key.lower() not in _HOP_BY_HOP:
                    self.send_header(key, value)
            self.send_header("Content-Length", str(len(payload)))
            self.end_headers()
            if self.command != "HEAD":
                self.wfile.write(payload)

        do_GET = do_POST = do_PUT = do_PATCH = do_DELETE = do_HEAD = \
            do_OPTIONS = _relay

    return Handler


@contextlib.contextmanager
def service_proxy(namespace, service, listen_host="127.0.0.1",
                  listen_port=None, verbose=False, access=None):
    """Run a background proxy to a Kubernetes service for the duration of
    the context; yields ``"host:port"``.

    Arguments:
        namespace (str): namespace of the service.
        service (str): ``[https:]service_name[:port_name]``.
        listen_host, listen_port: local bind address (a free port if None).
        verbose (bool): log every relayed request.
        access (KubeAccess): API-server credentials (default: the current
            kubectl context).
    """
    own = access is None
    if own:
        access = load_kube_access()
    prefix = "/api/v1/namespaces/{}/services/{}/proxy".format(namespace,
                                                              service)
    server = ThreadingHTTPServer((listen_host, listen_port or 0),
                                 _make_handler(access, prefix, verbose))
    server.daemon_threads = True
    thread = threading.Thread(target=server.serve_forever, daemon=True,
                              name="adaptdl-service-proxy")
    thread.start()
    try:
        yield "{}:{}".format(listen_host, server.server_address[1])
    finally:
        server.shutdown()
        server.server_close()
        thread.join(5)
        if own:
            access.close()
