set -x
( time python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2> gpurun_out/bench_full.time
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time
tail -3 gpurun_out/bench_full.time gpurun_out/bench_ref.time
cat gpurun_out/bench_full.json; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_full.err gpurun_out/bench_ref.err
nproc
