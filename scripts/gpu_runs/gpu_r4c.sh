# pass-1 phase accounting of the per-pass kernel on the mixed world (scripts/role_cycles.py, -DFSE_ROLE_CYCLES build)
mkdir -p gpurun_out
FSE_B200_LIB=_variants/libfse_role.so timeout 600 python scripts/role_cycles.py 8192 mixed > gpurun_out/r4c_role_mixed.txt 2> gpurun_out/r4c.err; cat gpurun_out/r4c_role_mixed.txt; tail -3 gpurun_out/r4c.err
FSE_B200_LIB=_variants/libfse_role.so timeout 600 python scripts/role_cycles.py 2048 water > gpurun_out/r4c_role_water.txt 2>> gpurun_out/r4c.err; cat gpurun_out/r4c_role_water.txt
