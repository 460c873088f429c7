#!/bin/sh
# One-GPU evidence run for profiles/ (developer tooling): GPU tests, the bench line, the ncu launch list of the
# same command, ncu --set full captures of the dominant kernels, the BASELINE config-4 sweep.
# usage (on the GPU box, from the repo root): sh tools/final_profiles.sh r02
R=${1:-r02}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/gputest_$R.log 2>&1; tail -2 $O/gputest_$R.log
python bench.py --steps 20 --warmup 5 > $O/bench_$R.json 2> $O/bench_$R.err; tail -c 400 $O/bench_$R.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_$R.csv \
    python bench.py --steps 2 --warmup 1 --no-c5 --no-cpu-baseline > /dev/null 2>&1
python tools/summarise_launches.py $O/launches_bench_$R.csv > $O/launches_bench_$R.txt
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:chol_lml_kernel -c 2 -o $O/chol_$R python tools/profile_kernels.py chol > /dev/null 2>&1
$NCU --cache-control none -k regex:chol_lml_kernel -c 2 -o $O/chol_nocacheflush_$R python tools/profile_kernels.py chol > /dev/null 2>&1
$NCU -k regex:gram_fused -c 2 -o $O/gram_$R python tools/profile_kernels.py chol > /dev/null 2>&1
$NCU -k regex:sweep_kernel -c 2 -o $O/sweep_$R python tools/profile_kernels.py sweep > /dev/null 2>&1
$NCU -k regex:mes_epilogue -c 2 -o $O/mes_$R python tools/profile_kernels.py mes > /dev/null 2>&1
{
  for f in chol chol_nocacheflush gram sweep mes; do
    echo "==== $O/${f}_$R.ncu-rep (ncu --set full --clock-control none$( [ $f = chol_nocacheflush ] && echo ' --cache-control none'))"
    python tools/ncu_summary.py $O/${f}_$R.ncu-rep | tail -21
  done
} > $O/ncu_kernels_$R.txt
python tools/ncu_lines.py $O/chol_$R.ncu-rep bayes-skopt_b200/libbgp.so chol_lml_kernelILi8ELi2E 40 > $O/chol_hot_lines_$R.txt 2>&1
python tools/ncu_lines.py $O/sweep_$R.ncu-rep bayes-skopt_b200/libbgp.so sweep_kernel 30 > $O/sweep_hot_lines_$R.txt 2>&1
python tools/lml_sweep.py 20,64,100,128,200,256,512,1024,2048,4096 64,256,1024 > $O/lml_sweep_$R.txt 2>&1
tail -12 $O/lml_sweep_$R.txt
