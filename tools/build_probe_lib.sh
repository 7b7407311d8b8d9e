#!/bin/bash
# Build a second copy of the library with -DKE_TUNING_PROBES (KE_PHASH_DBG, KE_JOIN_DEBUG, ... become live) for the probe
# tools: KE_LIB_PATH=kobato-eyes_b200/csrc/build/libkobato_probe.so python tools/probe_phash_roles.py
set -e
cd "$(dirname "$0")/../kobato-eyes_b200/csrc"
mkdir -p build/probe
pids=()
for f in ke_capi ke_multi ke_join ke_phash ke_ssim ke_synth ke_refine ke_resize_mma ke_scan ke_orb; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 -DKE_TUNING_PROBES \
        -I ../../include -I . -c -o build/probe/$f.o $f.cu &
    pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -gencode arch=compute_100a,code=sm_100a --shared -cudart static -o build/libkobato_probe.so build/probe/*.o
echo build/libkobato_probe.so
