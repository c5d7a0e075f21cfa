#!/usr/bin/env python
"""SASS evidence for ``profiles/sass/`` -- runs here, no GPU needed.

    python tools/sass_listing.py            # rewrite profiles/sass/*

* ``kernels.txt``: every kernel of ``libadl_b200.so`` with its register count,
  static shared memory and an opcode histogram of the instructions that say
  what the kernel is made of (tensor-core / TMA / multimem / system-scope
  memory operations, vector loads and stores, barriers, reductions);
* ``<name>.sass``: the full listing of the instantiations the headline
  workloads launch (see ``LISTED``).
"""

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adaptdl_b200", "_native", "libadl_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")

# file name -> demangled-name regex of the instantiation to list in full
LISTED = {
    "allreduce_twoshot_bf16_w8": r"allreduce_gns_kernel<__nv_bfloat16, 8, 0>",
    "allreduce_oneshot_bf16_w8": r"allreduce_oneshot_kernel<__nv_bfloat16, 8, 0>",
    "allreduce_nvls_bf16": r"allreduce_nvls_kernel<__nv_bfloat16, 0>",
    "allreduce_twoshot_f32_w8_adam": r"allreduce_gns_kernel<float, 8, 2>",
    "local_pair_bf16": r"local_kernel<__nv_bfloat16, 2, 0>",
    "finalize_stats": r"finalize_stats_kernel",
    "fused_optim_sgd_bf16_master": r"fused_optim_kernel<__nv_bfloat16, 0, true>",
    "fused_optim_adam_bf16_master": r"fused_optim_kernel<__nv_bfloat16, 1, true>",
    "bn_reduce_fwd_bf16": r"bn_reduce_kernel<__nv_bfloat16, false",
    "bn_apply_bwd_bf16": r"bn_apply_kernel<__nv_bfloat16, true",
    "ln_fwd_bf16_nv3": r"ln_fwd_kernel<__nv_bfloat16, 3>",
    "ln_bwd_bf16_nv3": r"ln_bwd_kernel<__nv_bfloat16, 3>",
    "ln_param_grad": r"ln_param_grad_kernel",
    "heads_split": r"heads_split_kernel",
    "colsum_bf16": r"colsum_kernel<__nv_bfloat16>",
    "gelu_dropout_bwd_bf16": r"gelu_dropout_bwd_kernel<__nv_bfloat16>",
    "slice_to_f32": r"slice_to_f32_kernel",
    "gemm_bias_act_bn256_2sm": r"gemm_bias_act_kernel<.*256.*2",
}

# opcode prefixes worth counting (everything else is arithmetic / control)
INTERESTING = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTCBAR", "LDTM",
               "STTM", "SYNCS", "LDGMC", "REDMC", "STGMC", "LDG", "STG",
               "LD.", "ST.", "ATOM", "RED", "MEMBAR", "ERRBAR", "CCTL",
               "BAR", "SHFL", "LDS", "STS", "FFMA2", "HFMA2", "FFMA",
               "CS2R", "S2UR", "NANOSLEEP", "ACQBULK", "UBLKCP")


def run(*cmd):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE,
                          stderr=subprocess.DEVNULL, text=True).stdout


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), check=True,
                         stdout=subprocess.PIPE, text=True).stdout
    return out.splitlines()


def main():
    if not os.path.exists(LIB):
        sys.exit("build the library first (python -c 'import "
                 "__graft_entry__ as g; g.build()')")
    os.makedirs(OUT, exist_ok=True)
    sass = run("cuobjdump", "-sass", LIB)
    usage = run("cuobjdump", "-res-usage", LIB)
    # resource usage: " Function NAME:\n  REG:.. STACK:.. SHARED:.. ..."
    res = {}
    current = None
    for line in usage.splitlines():
        m = re.match(r"\s*Function\s+(\S+):", line)
        if m:
            current = m.group(1)
            continue
        if current and "REG:" in line:
            res[current] = " ".join(
                tok for tok in line.split()
                if tok.split(":")[0] in ("REG", "STACK", "SHARED", "LOCAL"))
            current = None
    blocks, name, body = [], None, []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                blocks.append((name, body))
            name, body = m.group(1), []
        elif name:
            body.append(line)
    if name:
        blocks.append((name, body))
    pretty = dict(zip([b[0] for b in blocks],
                      demangle([b[0] for b in blocks])))
    op = re.compile(r"^\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    rows, totals = [], collections.Counter()
    for mangled, body in blocks:
        hist = collections.Counter()
        n = 0
        for line in body:
            m = op.match(line)
            if not m:
                continue
            n += 1
            code = m.group(1)
            if code.startswith(INTERESTING):
                hist[code] += 1
                totals[code.split(".")[0]] += 1
        rows.append((pretty[mangled], mangled, n, hist))
    rows.sort()
    with open(os.path.join(OUT, "kernels.txt"), "w") as f:
        f.write("# cuobjdump -sass / -res-usage of adaptdl_b200/_native/"
                "libadl_b200.so (sm_100a), one block per kernel:\n"
                "# demangled name | instructions | resources | selected "
                "opcodes with counts\n\n")
        for name, mangled, n, hist in rows:
            f.write("{}\n  {} instructions; {}\n".format(
                name, n, res.get(mangled, "?")))
            line = "  "
            for code, count in sorted(hist.items()):
                item = "{} x{}".format(code, count)
                if len(line) + len(item) > 100:
                    f.write(line.rstrip() + "\n")
                    line = "  "
                line += item + "  "
            if line.strip():
                f.write(line.rstrip() + "\n")
            f.write("\n")
        f.write("# opcode families over the whole library\n")
        for code, count in sorted(totals.items()):
            f.write("{} {}\n".format(code, count))
    # the Blackwell-only instructions, full mnemonics (tcgen05 = UTC*/LDTM,
    # TMA = UTMA*, mbarrier = SYNCS, multimem = LDGMC, packed fp32 = F*2)
    proof = collections.Counter(re.findall(
        r"\b((?:UTC|UTMA|LDTM|STTM|SYNCS|LDGMC|REDMC|FFMA2|FMUL2|FADD2|UCGABAR)"
        r"[A-Za-z0-9_.]*)", sass))
    with open(os.path.join(OUT, "mnemonics.txt"), "w") as f:
        for code, count in sorted(proof.items(), key=lambda kv: (-kv[1],
                                                                 kv[0])):
            f.write("{:7d} {}\n".format(count, code))
    listed = 0
    for fname, pattern in sorted(LISTED.items()):
        rx = re.compile(pattern)
        hit = next(((nm, mg) for nm, mg, _, _ in rows if rx.search(nm)),
                   None)
        if hit is None:
            print("no kernel matches", pattern)
            continue
        body = next(b for m, b in blocks if m == hit[1])
        with open(os.path.join(OUT, fname + ".sass"), "w") as f:
            f.write("// {}\n// {}\n".format(hit[0], res.get(hit[1], "")))
            # drop the encoding columns: mnemonic + operands is what a
            # reader needs, and the files stay reviewable
            for line in body:
                m = re.match(r"^\s*/\*([0-9a-f]{4})\*/\s+(.*?)\s*/\*", line)
                if m:
                    f.write("/*{}*/ {}\n".format(m.group(1), m.group(2)))
        listed += 1
    print("kernels: {}  listings: {}".format(len(rows), listed))


if __name__ == "__main__":
    main()
