#!/bin/bash
# Round evidence, run on the GPU box (gpurun): GPU tests, smoke, bench line, ncu launch list, one ncu --set full
# capture per kernel class.  Outputs land in gpurun_out/; tools/collect_profiles.py summarises them into profiles/.
set -u
if [ "${1:-all}" != "ncu" ]; then
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; tail -c 400 gpurun_out/bench.log; echo
fi
[ "${1:-all}" = "tests" ] && exit 0          # tests + smoke + bench only (kernels unchanged since the last ncu pass)
python tools/prof_detect.py > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 10 -c 10 --csv --log-file gpurun_out/launches.csv python tools/prof_detect.py > gpurun_out/ncu.log 2>&1
# kernel:launches-to-skip (tools/prof_detect.py runs 3 detections; take the 2nd one's launches: sep_a / sep_b run for
# octaves 1 and 2 -> skip 2 = octave 1 of the 2nd detection; fir_pass = the scalar passes of octave 3)
for k in oct0_mma:1 sep_a_mma:2 sep_b_mma:2 fir_pass:2 scan_tma:1 refine_kernel:1; do
  name=${k%%:*}; skip=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -o gpurun_out/prof_${name}_$skip \
      python tools/prof_detect.py > gpurun_out/ncu_$name.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
