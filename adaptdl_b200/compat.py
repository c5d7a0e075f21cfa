"""Import aliases: ``import adaptdl`` (and ``adaptdl_sched`` / ``adaptdl_ray``
/ ``adaptdl_cli``) resolve to this framework, so that training scripts
written against petuum/adaptdl run unmodified::

    import adaptdl
    import adaptdl.torch as adl            # -> adaptdl_b200.torch
    from adaptdl.torch.data import current_dataloader

The four top-level packages in the repository root (three lines each) call
:func:`alias`; nothing is copied or re-exported by hand -- a meta-path finder
maps every ``<alias>.x.y`` to ``<target>.x.y`` and registers the *same*
module object under both names (so module-level state, e.g. the current data
loader, is shared no matter which name imported it).
"""

import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import sys

__all__ = ["alias"]

_INSTALLED = {}


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):

    def __init__(self, name, target, renames=None):
        self.name, self.target = name, target
        # modules that are laid out differently here: alias-relative dotted
        # prefix -> target-relative dotted prefix ("" = the target itself),
        # longest prefix first
        self.renames = sorted((renames or {}).items(),
                              key=lambda kv: -len(kv[0]))

    def _real_name(self, fullname):
        if fullname == self.name:
            return self.target
        if not fullname.startswith(self.name + "."):
            return None
        rest = fullname[len(self.name) + 1:]
        for old, new in self.renames:
            if rest == old or rest.startswith(old + "."):
                rest = (new + rest[len(old):]).lstrip(".")
                break
        return self.target + ("." + rest if rest else "")

    def find_spec(self, fullname, path=None, target=None):
        real = self._real_name(fullname)
        if real is None:
            return None
        try:
            spec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            return None
        if spec is None:
            return None
        return importlib.machinery.ModuleSpec(
            fullname, self, origin=spec.origin,
            is_package=spec.submodule_search_locations is not None)

    def create_module(self, spec):
        # the real module object itself: one module, two names
        return importlib.import_module(self._real_name(spec.name))

    def exec_module(self, module):
        pass

    # ``python -m adaptdl_sched.allocator`` (the commands of the reference's
    # helm chart): runpy asks the loader for the code object and runs it as
    # ``__main__``
    def _real_loader(self, fullname):
        real = self._real_name(fullname)
        spec = importlib.util.find_spec(real)
        if spec is None or spec.loader is None:
            raise ImportError("no module named " + fullname)
        return real, spec

    def get_code(self, fullname):
        real, spec = self._real_loader(fullname)
        return spec.loader.get_code(real)

    def get_source(self, fullname):
        real, spec = self._real_loader(fullname)
        return spec.loader.get_source(real)

    def get_filename(self, fullname):
        return self._real_loader(fullname)[1].origin

    def is_package(self, fullname):
        spec = self._real_loader(fullname)[1]
        return spec.submodule_search_locations is not None


def alias(name, target, renames=None):
    """Make ``import <name>[.sub]`` return ``<target>[.sub]``; ``renames``
    maps sub-module paths of the alias that are named differently in the
    target."""
    if _INSTALLED.get(name) == target:
        return sys.modules.get(name)
    finder = _AliasFinder(name, target, renames)
    sys.meta_path.insert(0, finder)
    _INSTALLED[name] = target
    module = importlib.import_module(target)
    sys.modules[name] = module
    return module
