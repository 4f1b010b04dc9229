#!/bin/bash
# development helper: compile tuning variants of the library into build/variants/ (run here, they travel with gpurun);
# then `bash tools/sweep_variants.sh <workload>` on the GPU box times each one. One "name|nvcc -D flags" per line; a
# "_c<N>" / "_o<NN>" in the name makes the sweep set HWBRJ_PROBE_CTAS=N / HWBRJ_PROBE_CARVEOUT=NN for that run (k vs l:
# same small-ring kernel with the large L1 it allows and with the L1 of the 512-tuple ring -> drain size vs L1 size).
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
SRC=hwbloomradixjoin_b200/csrc/hwbrj.cu
spec=${1:-k2}
case $spec in
k2) list='
a_base|
b_claim_m4|-DHWBRJ_PROBE_CLAIM_AHEAD=1 -DHWBRJ_PROBE_MINBLOCKS=4
c_claim_ring256_m4|-DHWBRJ_PROBE_CLAIM_AHEAD=1 -DHWBRJ_PROBE_RING=256 -DHWBRJ_PROBE_MINBLOCKS=4
d_claim_ring128_m4|-DHWBRJ_PROBE_CLAIM_AHEAD=1 -DHWBRJ_PROBE_RING=128 -DHWBRJ_PROBE_MINBLOCKS=4
e_claim_ring128_m5_c5|-DHWBRJ_PROBE_CLAIM_AHEAD=1 -DHWBRJ_PROBE_RING=128 -DHWBRJ_PROBE_MINBLOCKS=5
f_ld_cg_all|-DHWBRJ_PROBE_LD=2
g_ld_ldg_all|-DHWBRJ_PROBE_LD=0
h_v8_m2_c2|-DHWBRJ_PROBE_V=8 -DHWBRJ_PROBE_MINBLOCKS=2
i_v8_m3_c3|-DHWBRJ_PROBE_V=8 -DHWBRJ_PROBE_MINBLOCKS=3
j_v2_m6_c6_claim_ring128|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=6 -DHWBRJ_PROBE_CLAIM_AHEAD=1 -DHWBRJ_PROBE_RING=128
k_ring128_m4|-DHWBRJ_PROBE_RING=128 -DHWBRJ_PROBE_MINBLOCKS=4
l_ring128_m4_o58|-DHWBRJ_PROBE_RING=128 -DHWBRJ_PROBE_MINBLOCKS=4
m_claim_ring128_m4_o58|-DHWBRJ_PROBE_CLAIM_AHEAD=1 -DHWBRJ_PROBE_RING=128 -DHWBRJ_PROBE_MINBLOCKS=4
' ;;
ablate) list='
a_base|
b_no_filter_loads|-DHWBRJ_K2_ABLATE=1
c_no_shared_cursor|-DHWBRJ_K2_ABLATE=2
d_stream_hash_only|-DHWBRJ_K2_ABLATE=3
' ;;
join) list='
a_base|
b_join_u4|-DHWBRJ_JOIN_UNROLL=4
c_join_u1|-DHWBRJ_JOIN_UNROLL=1
' ;;
r2b) list='
a_base|
b_join_book_per_batch|-DHWBRJ_JOIN_PENDING=0
' ;;
scatter) list='
a_base|
b_scatter_t512_m2|-DHWBRJ_SCATTER_THREADS=512 -DHWBRJ_SCATTER_MINBLOCKS=2
c_scatter_tile4096_t512_m2|-DHWBRJ_SCATTER_TILE=4096 -DHWBRJ_SCATTER_THREADS=512 -DHWBRJ_SCATTER_MINBLOCKS=2
d_scatter_tile1024_m5|-DHWBRJ_SCATTER_TILE=1024 -DHWBRJ_SCATTER_MINBLOCKS=5
e_scatter_stages3_m3|-DHWBRJ_SCATTER_STAGES=3 -DHWBRJ_SCATTER_MINBLOCKS=3
f_scatter_tile4096_t1024_m1|-DHWBRJ_SCATTER_TILE=4096 -DHWBRJ_SCATTER_THREADS=1024 -DHWBRJ_SCATTER_MINBLOCKS=1
' ;;
probe_shape) list='
a_base|
c_probe_w4_c8|-DHWBRJ_PROBE_WARPS=4
d_probe_w16_c2|-DHWBRJ_PROBE_WARPS=16
g_probe_w2_c16|-DHWBRJ_PROBE_WARPS=2
' ;;
schunk) list='
a_base|
b_schunk_64k|-DHWBRJ_JOIN_SCHUNK=65536
c_schunk_128k|-DHWBRJ_JOIN_SCHUNK=131072
d_schunk_16k|-DHWBRJ_JOIN_SCHUNK=16384
' ;;
k2grid) list='
a_base|
b_v2_m6_r256_c6|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=6 -DHWBRJ_PROBE_RING=256
c_v2_m5_r256_c5|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=5 -DHWBRJ_PROBE_RING=256
d_v2_m4_r512_c4|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=4
e_v3_m5_r256_c5|-DHWBRJ_PROBE_V=3 -DHWBRJ_PROBE_MINBLOCKS=5 -DHWBRJ_PROBE_RING=256
f_v3_m4_r512_c4|-DHWBRJ_PROBE_V=3 -DHWBRJ_PROBE_MINBLOCKS=4
g_v3_m4_r256_c4|-DHWBRJ_PROBE_V=3 -DHWBRJ_PROBE_MINBLOCKS=4 -DHWBRJ_PROBE_RING=256
h_v6_m3_r512_c3|-DHWBRJ_PROBE_V=6 -DHWBRJ_PROBE_MINBLOCKS=3
i_v5_m3_r512_c3|-DHWBRJ_PROBE_V=5 -DHWBRJ_PROBE_MINBLOCKS=3
j_v5_m4_r512_c4|-DHWBRJ_PROBE_V=5 -DHWBRJ_PROBE_MINBLOCKS=4
k_v3_m5_r512_c5|-DHWBRJ_PROBE_V=3 -DHWBRJ_PROBE_MINBLOCKS=5
' ;;
k2grid2) list='
a_base|
b_v2_m5_r256_c5|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=5 -DHWBRJ_PROBE_RING=256
c_v2_m5_r128_c5|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=5 -DHWBRJ_PROBE_RING=128
d_v2_m6_r128_c6|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=6 -DHWBRJ_PROBE_RING=128
e_v2_m5_r512_c5|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=5
f_v1_m8_r128_c8|-DHWBRJ_PROBE_V=1 -DHWBRJ_PROBE_MINBLOCKS=8 -DHWBRJ_PROBE_RING=128
g_v1_m7_r256_c7|-DHWBRJ_PROBE_V=1 -DHWBRJ_PROBE_MINBLOCKS=7 -DHWBRJ_PROBE_RING=256
h_v2_w16_m3_r256_c3|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_WARPS=16 -DHWBRJ_PROBE_MINBLOCKS=3 -DHWBRJ_PROBE_RING=256
i_v2_m5_r256_c5_o50|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=5 -DHWBRJ_PROBE_RING=256
j_v2_m5_r256_c5_o25|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=5 -DHWBRJ_PROBE_RING=256
k_v2_m5_r256_c4|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_MINBLOCKS=5 -DHWBRJ_PROBE_RING=256
l_v2_w4_m10_r256_c10|-DHWBRJ_PROBE_V=2 -DHWBRJ_PROBE_WARPS=4 -DHWBRJ_PROBE_MINBLOCKS=10 -DHWBRJ_PROBE_RING=256
' ;;
k2shape) list='
a_base|
b_v4_m4_r512|-DHWBRJ_PROBE_V=4 -DHWBRJ_PROBE_MINBLOCKS=4 -DHWBRJ_PROBE_RING=512
' ;;
*) echo "unknown spec $spec"; exit 1 ;;
esac
rm -f build/variants/lib_*.so
while IFS='|' read -r name flags; do
  [ -z "$name" ] && continue
  nvcc $F $flags -o build/variants/lib_$name.so $SRC -ldl &
done <<< "$list"
wait
for f in build/variants/lib_*.so; do
  echo "$(basename $f) $(cuobjdump --dump-resource-usage $f | grep -A1 'k_probe_compactILi6E' | grep -o 'REG:[0-9]* STACK:[0-9]*')"
done
