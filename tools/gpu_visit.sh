#!/bin/bash
# One parametrised GPU visit (replaces the per-visit scripts of round 1).
#   usage: tools/gpu_visit.sh <tag> <step> [<step> ...]      outputs: gpurun_out/<tag>_*
# steps: probe | tests_dycore | tests_physics | tests | variants | bench | bench_ref | launches | prof_stage | prof_kessler |
#        physics | smoke | config3 | multi | benchN | config3N (NGPU=2,4,8) | config3strong1
# Env: VARIANTS (for `variants`, default "5 4"), BENCH_ARGS, PROF_KERNEL (regex for prof_stage, default k_stage)
tag=$1; shift
mkdir -p gpurun_out
o=gpurun_out/$tag
for step in "$@"; do
  echo "=== $step"
  case $step in
    probe)         ./tools/bin/issue_probe > ${o}_issue_probe.jsonl 2>&1; tail -2 ${o}_issue_probe.jsonl ;;
    tests_dycore)  timeout 900 python -m pytest tests/test_gpu_dycore.py -m gpu -q -x --timeout 300 2>&1 | tail -8 | tee ${o}_tests_dycore.log ;;
    tests_physics) timeout 900 python -m pytest tests/test_gpu_physics.py -m gpu -q -x --timeout 300 2>&1 | tail -8 | tee ${o}_tests_physics.log ;;
    tests)         timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 | tee ${o}_tests.log ;;
    variants)      for v in ${VARIANTS:-5 4}; do echo "-- MW_TILE_VARIANT=$v"; MW_TILE_VARIANT=$v timeout 600 python tools/probe_dycore.py 2>&1 | tail -3; done | tee ${o}_variants.log ;;
    bench)         timeout 900 python bench.py $BENCH_ARGS 2>${o}_bench.err | tail -1 | tee ${o}_bench_n1.json ;;
    bench_ref)     timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>${o}_bench.err | tail -1 | tee ${o}_bench_ref.json ;;
    config3)       timeout 900 python bench.py --workload config3 --steps 5 --warmup 2 --no-cpu-baseline 2>>${o}_bench.err | tail -1 | tee ${o}_bench_config3_n1.json ;;
    launches)      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${o}_launches.csv \
                     python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > ${o}_launches_bench.log 2>&1; tail -3 ${o}_launches_bench.log ;;
    prof_stage)    timeout 1500 ncu --set full --clock-control none --import-source on -k regex:${PROF_KERNEL:-k_stage} -s 3 -c 1 -f -o ${o}_prof_stage \
                     python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 0 > ${o}_prof_bench.log 2>&1; tail -2 ${o}_prof_bench.log ;;
    prof_kessler)  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_kessler_single -s 1 -c 1 -f -o ${o}_prof_kessler \
                     python tools/probe_physics.py > ${o}_prof_kessler.log 2>&1; tail -2 ${o}_prof_kessler.log ;;
    physics)       timeout 600 python tools/probe_physics.py 2>&1 | tail -6 | tee ${o}_physics_probe.jsonl ;;
    smoke)         timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee ${o}_smoke.log ;;
    multi)         timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_host_driver.py -m gpu -q --timeout 900 2>&1 | tail -12 | tee ${o}_tests_multi.log ;;
    benchN)        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29511 \
                     bench.py --gpus $NGPU --steps 10 --warmup 3 2>>${o}_bench.err | tail -1 | tee ${o}_bench_n$NGPU.json ;;
    config3N)      for sc in strong weak; do timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port 29512 \
                     bench.py --gpus $NGPU --workload config3 --scaling $sc --steps 5 --warmup 3 --no-cpu-baseline 2>>${o}_bench.err | tail -1 | tee ${o}_bench_config3_${sc}_n$NGPU.json; done ;;
    config3strong1) timeout 900 python bench.py --workload config3 --scaling strong --steps 3 --warmup 3 --no-cpu-baseline 2>>${o}_bench.err | tail -1 | tee ${o}_bench_config3_strong_n1.json ;;
    *) echo "unknown step $step" ;;
  esac
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
