# chain kernel (pass 1 -> pass 2 in one CTA): parity on the split-test world, then the default bench line with and without it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "chain or split_modes" > gpurun_out/r4a_chain_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4a_chain_tests.log; tail -5 gpurun_out/r4a_chain_tests.log
for cfg in "0 1" "1 1" "1 0" "1 2"; do
set -- $cfg
FSE_CHAIN=$1 FSE_P2_SPLIT=$2 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r4a_chain$1_split$2.json 2> gpurun_out/r4a.err
python -c "
import json
d=json.loads(open('gpurun_out/r4a_chain$1_split$2.json').read().strip().splitlines()[-1]); print('chain $1 split $2', round(d['value'],3), round(d['ms_per_step'],3), d['state']['hash'], round(d['e2e']['ms_per_step'],2)); print(d['roofline'].get('phase_ms_by_iteration'))"
done
