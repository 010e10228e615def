#!/bin/bash
# ncu --set full captures of the hot kernels (one launch each, after a plain run of the same command).
# usage: tools/profile_kernels.sh <tag> [names...]     (run under gpurun; reports land in gpurun_out/)
set -u
TAG=${1:-r1}
shift || true
ONLY="$*"
OUT=gpurun_out
mkdir -p $OUT
prof () {  # name kernel-regex microbench-args...
  local name=$1; local regex=$2; shift 2
  if [ -n "$ONLY" ] && ! echo " $ONLY " | grep -q " $name "; then return; fi
  python tools/microbench.py "$@" > $OUT/plain_${name}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s 3 -c 1 \
      -o $OUT/${name}_${TAG} -f python tools/microbench.py "$@" > $OUT/ncu_${name}.log 2>&1
  tail -1 $OUT/plain_${name}.log
}
if [ "${TAG#r1}" != "$TAG" ]; then     # round-1 kernel names
prof corr_fwd corr_fwd_md4 corr_fwd --shapes 64x32x96x128
prof corr_bwd corr_bwd_md4 corr_bwd --shapes 64x32x96x128
prof warp_fwd warp_fwd_kernel warp --flow smooth --shapes 64x32x96x128
prof census_fwd census_fwd_kernel census --shapes 8x3x384x512
prof census_bwd census_bwd_kernel census --shapes 8x3x384x512
else
prof corr_fwd corr_fwd_md4_p2 corr_fwd --shapes 64x32x96x128
prof corr_fwd_cfg2 corr_fwd_md4_p2 corr_fwd --shapes 16x32x96x128
prof corr_bwd corr_bwd_md4 corr_bwd --shapes 64x32x96x128
prof corr_bwd_cfg2 corr_bwd_md4 corr_bwd --shapes 16x32x96x128
prof warp_fwd warp_fwd_kernel warp --flow smooth --shapes 16x32x96x128
prof warp_bwd_gx warp_bwd_lean warp --flow smooth --shapes 16x32x96x128
prof census_fwd census_fwd_sym census --shapes 8x3x384x512
prof census_fwd_b16 census_fwd_sym census --shapes 16x3x384x512
prof census_bwd_b16 census_bwd_sym census --shapes 16x3x384x512
prof census_bwd census_bwd_sym census --shapes 8x3x384x512
fi
prof stencil_fwd 'stencil_mv_kernel' stencil
prof stencil_bwd 'stencil_mv_bwd_kernel' stencil
prof trisolve trisolve_scan_kernel stencil
ls -la $OUT/*_${TAG}.ncu-rep
