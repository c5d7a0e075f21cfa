"""Stand-in for the `portpicker` package (not installable offline): the
reference only calls pick_unused_port()."""
import socket


def pick_unused_port():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("", 0))
        return s.getsockname()[1]
