#!/bin/bash
# tuning builds of the same ABI: scripts/build_variants.sh "W C [extra -D flags]" ...  -> build/var/libfrx_<tag>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build/var
for spec in "$@"; do
  set -- $spec
  W=$1; C=$2; shift 2; EXTRA="$*"
  tag="W${W}_C${C}$(echo "$EXTRA" | tr -d ' =-' | sed 's/DFRX_/_/g')"
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -shared -I include \
    -DFRX_WARPS_PER_CTA=$W -DFRX_MIN_CTAS=$C $EXTRA -o build/var/libfrx_${tag}.so \
    frenetix_motion_planner_b200/csrc/frx_kernels.cu frenetix_motion_planner_b200/csrc/frx_capi.cu &
done
wait
ls -la build/var
