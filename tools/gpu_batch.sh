#!/bin/bash
# One gpurun call = several experiments; every step under its own timeout, logs under gpurun_out/$TAG/
TAG=${1:-batch}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { name=$1; shift; echo "=== $name: $*" | tee -a $OUT/summary.txt; ( timeout -s KILL ${TMO:-300} "$@" ) > $OUT/$name.log 2>&1; echo "rc=$? ($name)" | tee -a $OUT/summary.txt; tail -${TAILN:-15} $OUT/$name.log | tee -a $OUT/summary.txt; }
make -C sober_b200/csrc -j16 > $OUT/make.log 2>&1 || { echo "BUILD FAILED"; tail -20 $OUT/make.log; }
for step in "$@"; do
  case $step in
    carpanel) run carpanel_tests python -m pytest tests/test_car_panel.py -x -q -m gpu ;;
    timecar)  TAILN=40 run time_car python tools/time_car.py ;;
    k1var)    for lib in k1_variants/*.so; do SOBER_B200_LIB=$lib TMO=120 TAILN=2 run k1_$(basename $lib .so) python tools/k1_time.py; done ;;
    stage5)   TAILN=25 run stage_c5 python tools/stage_breakdown.py c5 ;;
    stage2)   TAILN=25 run stage_c2 python tools/stage_breakdown.py c2 ;;
    nys5)     TAILN=40 run nystrom_c5 python tools/nystrom_breakdown.py 2000 999 ;;
    tests)    TMO=900 run pytest_gpu python -m pytest tests -x -q -m gpu ;;
    smoke)    run smoke python __graft_entry__.py smoke ;;
    ncuk1)    TMO=600 run ncu_k1 ncu --set full --clock-control none --import-source on -k regex:group_records -s 4 -c 1 -f -o $OUT/k1prof python tools/k1_time.py ;;
    ncusmoke) TMO=600 TAILN=60 run ncu_smoke ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" ;;
    parbrk)   TAILN=20 run parity_breakdown python tools/parity_breakdown.py ;;
    ncucar)   TMO=600 TAILN=5 run ncu_car ncu --metrics gpu__time_duration.sum --clock-control none -k regex:car_panel -c 400 --csv --log-file $OUT/launches_car.csv python tools/time_car.py ;;
    shard2)   TMO=600 run sharded_tests python -m pytest tests/test_gpu_sharded.py -x -q -m gpu ;;
    mbench2)  TMO=900 TAILN=3 run bench_2gpu python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ;;
    mbench4)  TMO=900 TAILN=3 run bench_4gpu python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 5 --warmup 3 ;;
    mbench8)  TMO=900 TAILN=3 run bench_8gpu python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 ;;
    projbrk)  TAILN=30 run projector_breakdown python tools/projector_breakdown.py ;;
    bitsmma)  TMO=240 TAILN=25 run bits_mma_tests python -m pytest tests/test_bits_mma.py -x -q -m gpu ;;
    bench4)   TMO=600 TAILN=3 run bench_c4 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline ;;
    bench3)   TMO=600 TAILN=3 run bench_c3 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline ;;
    bitstime) TMO=300 TAILN=6 run bits_time python tools/bits_time.py ;;
    ncubits)  TMO=600 run ncu_bits ncu --set full --clock-control none --import-source on -k regex:group_bits_mma -s 2 -c 1 -f -o $OUT/bitsprof python tools/bits_time.py ;;
    sanit)    TMO=900 TAILN=12 run sanitizer_racecheck compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_car_panel.py tests/test_bits_mma.py -x -q -m gpu -k "400-200-0 or 130-61-64 or early_stop or 256-0.1" ;
              TMO=900 TAILN=12 run sanitizer_memcheck compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_car_panel.py tests/test_bits_mma.py tests/test_pi.py -x -q -m gpu -k "400-200-0 or 1000-500-0 or early_stop or 256-0.1 or 12_345 or pi_rbf" ;;
    sanitbits) TMO=900 TAILN=30 run sanitizer_racecheck_bits compute-sanitizer --tool racecheck --print-limit 40 python -m pytest tests/test_bits_mma.py -x -q -m gpu -k "256-0.1 or 4_100 or 12_345" ;;
    sanitall) TMO=1500 TAILN=15 run sanitizer_memcheck_suite compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests -x -q -m gpu -k "not full_size and not pipelined and not sharded and not at_size and not car_cluster and not cholesky" ;;
    ncustep)  TMO=900 TAILN=3 run ncu_step ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_c2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras ;;
    bench1)   TMO=600 TAILN=3 run bench_c1 python bench.py --workload c1 --steps 10 --warmup 3 ;;
    stagepar) TAILN=25 run stage_c2_parity python tools/stage_breakdown.py c2 parity ;;
    bench2)   TMO=600 TAILN=3 run bench_c2 python bench.py --steps 10 --warmup 3 ;;
    bench5)   TMO=600 TAILN=3 run bench_c5 python bench.py --workload c5 --steps 5 --warmup 2 --no-cpu-baseline ;;
    *)        echo "unknown step $step" ;;
  esac
done
