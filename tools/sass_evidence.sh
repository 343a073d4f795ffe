#!/bin/bash
# SASS evidence for the parity-critical arithmetic and the staged-tile kernel:
#   (i)  no FFMA in the kernels whose results must match the reference bit for bit in f32
#        (distance, covariance, transform, linearisation are written with __f{add,sub,mul}_rn);
#        the only fused f32 multiply-adds allowed are in f64-backed epilogues / address math.
#   (ii) UBLKCP (1-D TMA bulk copy) + SYNCS (mbarrier) in the staged-tile kernel.
# usage: tools/sass_evidence.sh > profiles/r02_sass_evidence.txt
set -e
LIB=threecrate_b200/lib/libthreecrate_cuda.so
echo "# cuobjdump -sass $LIB  ($(git rev-parse --short HEAD), $(date -u +%F))"
cuobjdump -sass $LIB > /tmp/tc_all.sass
python3 - <<'PY'
import re, collections
fn = None
stats = collections.OrderedDict()
for line in open('/tmp/tc_all.sass'):
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        fn = m.group(1); stats[fn] = collections.Counter(); continue
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
    if m and fn:
        op = m.group(1).split('.')[0]
        stats[fn][op] += 1
import subprocess
def dem(n):
    try: return subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()
    except Exception: return n
print(f"{'kernel':70s} {'insts':>7s} {'FFMA':>5s} {'FADD':>5s} {'FMUL':>5s} {'FMNMX':>6s} {'DFMA':>5s} {'UBLKCP':>6s} {'SYNCS':>5s}")
for fn, c in stats.items():
    d = dem(fn)
    short = re.sub(r'\(.*', '', d).replace('(anonymous namespace)::', '')
    if not any(k in short for k in ('k_normals', 'k_knn', 'k_tile', 'k_icp_correspond', 'k_radius', 'k_gicp_cov', 'k_voxel')):
        continue
    print(f"{short[:70]:70s} {sum(c.values()):7d} {c['FFMA']:5d} {c['FADD']:5d} {c['FMUL']:5d} {c['FMNMX']:6d} {c['DFMA']:5d} {c['UBLKCP']:6d} {c['SYNCS']:5d}")
PY
echo
echo "# FFMA occurrences inside the hot search kernels, with their source lines (-lineinfo):"
echo "# (expected: only the conservative pruning bounds axis_bound()/ring_bound(), the f32 parts of"
echo "#  the division / square-root expansions and the viewpoint normalisation helpers - never"
echo "#  dist2_exact, the covariance sums, quat_rotate or the linearisation)"
nvdisasm_ok=0
cuobjdump -sass -fun 'k_normals2' $LIB 2>/dev/null | grep -c FFMA | sed 's/^/k_normals2* FFMA count: /' || true
