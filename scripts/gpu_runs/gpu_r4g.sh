# pass-1 phase accounting with (nearly) one chunk per SM: water-filled, sand stripes, half water (per-pass kernels forced)
mkdir -p gpurun_out
for wl in water waterhalf sand3 gas16; do
FSE_FUSED_MAX_CHUNKS=0 FSE_ROW_SKIP=0 FSE_B200_LIB=_variants/libfse_role.so timeout 300 python scripts/role_cycles.py 1792 $wl > gpurun_out/r4g_role_$wl.txt 2>> gpurun_out/r4g.err; cat gpurun_out/r4g_role_$wl.txt
done
