#!/usr/bin/env python
"""Documentation consistency check (CPU only, no dependencies).

For every tracked markdown file:
* relative links ``[text](path)`` must point at something that exists;
* back-ticked repository paths (``adaptdl_b200/...``, ``csrc/...``,
  ``tools/...``, ``tests/...``, ``docs/...``, ``deploy/...``,
  ``examples/...``, ``tutorial/...``, ``profiles/...``) must exist
  (``path:line`` and ``path::test`` suffixes are stripped, globs expanded);
* fenced ``python`` blocks must compile.

Exit code 1 and one line per problem otherwise. Used by
``.github/workflows/docs.yaml`` and ``tests/test_tools.py``.
"""
import glob
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOP_LEVEL = ("adaptdl_b200", "csrc", "tools", "tests", "docs", "deploy",
             "examples", "tutorial", "profiles", "baseline")
# created by the offline reference install, not tracked
GENERATED = ("baseline/_ref",)
# files that quote paths of OTHER trees (the reference, retrieved snippets)
SKIP = {"SURVEY.md", "PAPERS.md", "SNIPPETS.md", "BASELINE.md", "VERDICT.md",
        "ADVICE.md"}

LINK = re.compile(r"\[[^\]]*\]\(([^)#\s]+)(?:#[^)]*)?\)")
TICKED = re.compile(r"`([^`\s]+)`")
FENCE = re.compile(r"^```(\w*)\s*$")


def markdown_files():
    found = []
    for base, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (
            ".git", "gpurun_out", "_ref", "__pycache__", ".pytest_cache",
            "build")]
        for name in files:
            if name.endswith(".md") and name not in SKIP:
                found.append(os.path.join(base, name))
    return sorted(found)


def repo_path(token):
    """The repository path a back-ticked token refers to, or ``None``."""
    token = token.split("::")[0]
    token = re.sub(r":\d+(-\d+)?$", "", token).rstrip(".,;:")
    if token.startswith("./"):
        token = token[2:]
    head = token.split("/")[0]
    if "/" not in token or head not in TOP_LEVEL:
        return None
    if token.startswith(GENERATED):       # git-ignored, absent from a clone
        return None
    if any(ch in token for ch in "<>{}$|") or "..." in token or \
            "…" in token:
        return None
    return token


def exists(path):
    full = os.path.join(ROOT, path)
    if any(ch in path for ch in "*?["):
        return bool(glob.glob(full))
    return os.path.exists(full)


def check_file(path):
    problems = []
    rel = os.path.relpath(path, ROOT)
    here = os.path.dirname(path)
    block, block_lang, block_start = None, "", 0
    with open(path, encoding="utf-8") as f:
        lines = f.read().split("\n")
    for number, line in enumerate(lines, 1):
        fence = FENCE.match(line.strip())
        if fence:
            if block is None:
                block, block_lang, block_start = [], fence.group(1), number
            else:
                if block_lang in ("python", "py"):
                    try:
                        compile("\n".join(block), rel, "exec")
                    except SyntaxError as exc:
                        problems.append("{}:{}: python block does not "
                                        "compile: {}".format(
                                            rel, block_start + (exc.lineno
                                                                or 0),
                                            exc.msg))
                block = None
            continue
        if block is not None:
            block.append(line)
            continue
        for target in LINK.findall(line):
            if "://" in target or target.startswith("mailto:"):
                continue
            if not os.path.exists(os.path.normpath(
                    os.path.join(here, target))) and not exists(target):
                problems.append("{}:{}: broken link {}".format(
                    rel, number, target))
        for token in TICKED.findall(line):
            candidate = repo_path(token)
            if candidate is not None and not exists(candidate):
                problems.append("{}:{}: no such path `{}`".format(
                    rel, number, candidate))
    return problems


def main():
    problems = []
    for path in markdown_files():
        problems.extend(check_file(path))
    for line in problems:
        print(line)
    print("{} markdown files, {} problems".format(len(markdown_files()),
                                                  len(problems)))
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
