#!/usr/bin/env python
"""SASS evidence (VERDICT r1 'missing #8'):
  (i)  every FFMA of the search / tile / ICP translation units, attributed to its CUDA source line
       (nvdisasm -g): parity needs the reference's unfused f32 arithmetic in dist2_exact, the
       centroid / covariance sums, quat_rotate and the ICP linearisation, so those lines must not
       appear; FFMA is legitimate only inside the IEEE division / square-root expansions
       (__fdiv_rn, __fsqrt_rn), the conservative pruning bounds and other non-parity code.
  (ii) the TMA / mbarrier mnemonics of the staged-tile kernel: UBLKCP (cp.async.bulk) and SYNCS.
usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "threecrate_b200", "lib", "libthreecrate_cuda.so")
PARITY_LINES = {  # source lines that must never hold an FFMA: file -> regex on the source text
    "dist2_exact|xadd|xsub|xmul": None,
}


def main():
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True,
                          text=True).stdout.strip()
    print(f"# SASS evidence for {os.path.relpath(LIB, ROOT)} at commit {head} (nvdisasm -g per cubin)")
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=td, capture_output=True)
        for unit in ("tc_search", "tc_tile", "tc_icp"):
            cub = os.path.join(td, f"{unit}.sm_100a.cubin")
            out = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
            cur, per_line, ops = None, collections.Counter(), collections.Counter()
            for ln in out.splitlines():
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                m = re.match(r"\s+(?:/\*[0-9a-f]+\*/\s+)?(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
                if not m:
                    continue
                op = m.group(1)
                ops[op] += 1
                if op == "FFMA":
                    per_line[cur] += 1
            print(f"\n## {unit}.cu: {sum(ops.values())} SASS instructions; FFMA {ops['FFMA']}, FADD "
                  f"{ops['FADD']}, FMUL {ops['FMUL']}, FMNMX {ops['FMNMX']}, DFMA {ops['DFMA']}, "
                  f"UBLKCP {ops['UBLKCP']}, SYNCS {ops['SYNCS']}")
            print("FFMA by source line:")
            for (f, l), c in sorted(per_line.items(), key=lambda kv: -kv[1]):
                src = ""
                for base in (os.path.join(ROOT, "threecrate_b200", "csrc"),):
                    p = os.path.join(base, f)
                    if os.path.exists(p):
                        src = open(p).read().splitlines()[l - 1].strip()
                print(f"  {c:5d}  {f}:{l}  {src[:100]}")
            bad = [k for k in per_line if k[0] == "tc_internal.cuh" and any(
                t in open(os.path.join(ROOT, "threecrate_b200", "csrc", "tc_internal.cuh")).read().splitlines()[k[1] - 1]
                for t in ("__fadd_rn", "__fsub_rn", "__fmul_rn", "dist2_exact"))]
            print("parity-critical helpers (xadd/xsub/xmul/dist2_exact) holding an FFMA:", bad or "none")


if __name__ == "__main__":
    main()
