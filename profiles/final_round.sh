#!/bin/bash
# End-of-round evidence in one GPU lease (outputs under gpurun_out/<tag>/, summaries are copied into profiles/ afterwards):
#   bash profiles/final_round.sh r2_v
tag=${1:-r2_v}
out=gpurun_out/$tag
mkdir -p $out
t0=$(date +%s)
python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -1 $out/pytest_gpu.log) $(( $(date +%s) - t0 )) s"
python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "smoke rc=$? $(( $(date +%s) - t0 )) s"
python bench.py > $out/bench_8k_photo.json 2> $out/bench_8k_photo.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"
python bench.py --kind random --no-reference-gpu --no-cpu-baseline > $out/bench_8k_random.json 2>/dev/null
python bench.py --subsampling 4:2:0 --interleaved 1 --no-reference-gpu --no-cpu-baseline > $out/bench_8k_photo_420_interleaved.json 2>/dev/null
GPUJPEG_B200_STRIPES=1 python bench.py --no-reference-gpu --no-cpu-baseline > $out/bench_8k_photo_no_stripes.json 2>/dev/null
GPUJPEG_B200_PDL=0 python bench.py --no-reference-gpu --no-cpu-baseline > $out/bench_8k_photo_no_pdl.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_arm.json 2>/dev/null; echo "bench variants done $(( $(date +%s) - t0 )) s"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_8k_photo.csv \
    python bench.py --steps 5 --warmup 3 --e2e-workers 1 --no-reference-gpu --no-cpu-baseline > $out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 12 -c 6 -f -o $out/prof python profiles/run_step.py > $out/prof.log 2>&1
echo "ncu done $(( $(date +%s) - t0 )) s"
{ compute-sanitizer --tool memcheck python profiles/sanitize_cases.py; compute-sanitizer --tool racecheck python profiles/sanitize_cases.py quick; } > $out/sanitizer.txt 2>&1
echo "sanitizer done $(( $(date +%s) - t0 )) s"; grep -c "^ok" $out/sanitizer.txt; grep "SUMMARY" $out/sanitizer.txt
head -c 600 $out/bench_8k_photo.json; echo
