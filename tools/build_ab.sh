#!/bin/bash
# Builds a copy of the current csrc/ into ab/libpm_<name>.so (A/B and profiling builds; development aid).
#   tools/build_ab.sh prof -DPM_RANSAC_PROFILE      (extra flags go to ransac.cu)
#   NVEXTRA=-DPM_I8_EPI=0 tools/build_ab.sh epi0    (NVEXTRA goes to every file)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
rm -rf /tmp/abuild && mkdir -p /tmp/abuild/include ab
cp -r reconstructor_b200/csrc /tmp/abuild/csrc && cp include/pairmatch_b200.h /tmp/abuild/include/
cd /tmp/abuild/csrc && rm -rf build
sed -i 's#\.\./\.\./include#../include#' Makefile api.cu
if [ -n "$*" ]; then sed -i "s/^EXTRA_ransac := -fmad=false/EXTRA_ransac := -fmad=false $*/" Makefile; fi
make -j8 NVEXTRA="$NVEXTRA" OUT=/root/repo/ab/libpm_$name.so 2>&1 | grep -i "error" || true
ls -la /root/repo/ab/libpm_$name.so
