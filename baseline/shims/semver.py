"""Stand-in for the `semver` package: the reference uses
VersionInfo.isvalid / VersionInfo.parse(...).major only."""
import re

_RE = re.compile(r"^(\d+)\.(\d+)\.(\d+)(?:[-+].*)?$")


class VersionInfo(object):
    def __init__(self, major, minor=0, patch=0):
        self.major, self.minor, self.patch = major, minor, patch

    @staticmethod
    def isvalid(version):
        return bool(version) and _RE.match(str(version)) is not None

    @staticmethod
    def parse(version):
        m = _RE.match(str(version))
        if not m:
            raise ValueError(version)
        return VersionInfo(*(int(g) for g in m.groups()))
