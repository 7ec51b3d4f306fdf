#!/bin/bash
# One-GPU evidence pass on the GPU box (run through gpurun): tests, smoke, bench lines, the ncu
# launch list and one full capture of the raster kernel, the config-5 sweep, a fuzz soak.
#   gpurun --timeout 1800 -- 'R=r02 bash tools/collect_evidence.sh'
# Outputs land in gpurun_out/; copy what should be judged into profiles/.
R=${R:-r02}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${R}_pytest_gpu.log 2>&1; tail -2 $O/${R}_pytest_gpu.log
python __graft_entry__.py smoke > $O/${R}_smoke.log 2>&1; tail -1 $O/${R}_smoke.log
python bench.py > $O/${R}_bench_c3_1gpu.json 2> $O/${R}_bench.err; cut -c1-200 $O/${R}_bench_c3_1gpu.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2>> $O/${R}_bench.err
cut -c1-200 $O/${R}_bench_reference.json
# profiler passes: only after the plain runs above exited; numbers printed under ncu are not bench values
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-reference-gpu --no-config4 > $O/${R}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 4 -c 1 -f -o $O/${R}_raster_full \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-reference-gpu --no-config4 > $O/${R}_ncu_full.log 2>&1
ls -la $O/${R}_raster_full.ncu-rep
python tools/time_small_batch.py > $O/${R}_small_batch.txt 2>&1; tail -4 $O/${R}_small_batch.txt
python tools/time_engine.py > $O/${R}_engine_rates.txt 2>&1; cat $O/${R}_engine_rates.txt
if [ "${SWEEP:-1}" = "1" ]; then NCU=1 STEPS=5 bash tools/sweep.sh > $O/${R}_sweep.txt 2>&1; tail -3 $O/${R}_sweep.txt; fi
GGS_FUZZ_TRIALS=${FUZZ:-400} python -m pytest tests/test_gpu_parity.py -q -k randomised > $O/${R}_fuzz.log 2>&1; tail -1 $O/${R}_fuzz.log
python tools/reference_gpu_compare.py --out $O/${R}_reference_gpu_compare.json > $O/${R}_reference_gpu_compare.log 2>&1; tail -3 $O/${R}_reference_gpu_compare.log
python tools/time_sa.py > $O/${R}_search_rates.log 2>&1
TQDM_DISABLE=1 python examples/run_ga_synthetic.py --side 256 --splats 512 --pop 32 --generations 30000 >> $O/${R}_search_rates.log 2>&1
TQDM_DISABLE=1 python examples/run_ga_synthetic.py --side 128 --splats 100 --pop 32 --generations 30000 >> $O/${R}_search_rates.log 2>&1
TQDM_DISABLE=1 python examples/run_ga_synthetic.py --side 256 --splats 1000 --pop 1024 --generations 300 >> $O/${R}_search_rates.log 2>&1
cat $O/${R}_search_rates.log
