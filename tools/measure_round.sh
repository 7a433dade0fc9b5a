#!/usr/bin/env bash
# ONE GPU-box call that re-measures the single-GPU files under profiles/ (see profiles/README.md for what each one is):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/measure_round.sh r3'
# Everything is written to gpurun_out/<tag>_*; copy what should be judged into profiles/.  Bench values come from plain runs;
# the ncu passes run afterwards, each only after its own command has exited 0 without ncu, and their timings are never quoted.
set -uo pipefail
cd "$(dirname "$0")/.."
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p "$OUT"
run() { echo "== $*" >&2; timeout "${T:-900}" "$@"; }

# 1. parity first: the GPU test-suite and the smoke check
T=1500 run python -m pytest tests -m gpu -x -q > "$OUT/${TAG}_pytest_gpu.log" 2>&1 || { tail -30 "$OUT/${TAG}_pytest_gpu.log"; exit 1; }
run python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/${TAG}_smoke.log" 2>&1 || exit 1

# 2. bench lines of the four precision paths and the reference arm (plain runs)
run python bench.py --warmup 3 --steps 100 > "$OUT/bench_${TAG}_fp16e4m3.json" 2> "$OUT/${TAG}_bench.err" || exit 1
for p in bf16x3 fp16x3; do
  run python bench.py --warmup 3 --steps 100 --precision $p --no-cpu-baseline > "$OUT/bench_${TAG}_$p.json" 2>> "$OUT/${TAG}_bench.err"
done
run python bench.py --warmup 3 --steps 30 --precision fp32 --no-cpu-baseline > "$OUT/bench_${TAG}_fp32.json" 2>> "$OUT/${TAG}_bench.err"
run python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_${TAG}_reference.json" 2>> "$OUT/${TAG}_bench.err"

# 3. BASELINE configs 3 / 4 / 5 on one GPU, and the copy ceiling
run python tools/bench_configs.py fp16e4m3 > "$OUT/${TAG}_configs_fp16e4m3.json" 2>> "$OUT/${TAG}_bench.err"
run python tools/grid_search.py 14 fp16e4m3 > "$OUT/${TAG}_grid.json" 2>> "$OUT/${TAG}_bench.err"
run python tools/mcmc_bench.py 100000 1000 fp16e4m3 > "$OUT/${TAG}_mcmc.json" 2>> "$OUT/${TAG}_bench.err"
run python tools/train_bench.py 5 > "$OUT/${TAG}_train.json" 2>> "$OUT/${TAG}_bench.err"
run python tools/pcie_probe.py --gpus 1 > "$OUT/${TAG}_pcie_probe_n1.json" 2>> "$OUT/${TAG}_bench.err"
[[ -x tools/ffma2_probe ]] && run ./tools/ffma2_probe > "$OUT/${TAG}_ffma2_probe.log" 2>&1
[[ -f tools/ab/libvae21_mid.so && -f tools/ab/libvae21_xpf.so ]] && \
  run python tools/fp32_ab.py mid=tools/ab/libvae21_mid.so xpf=tools/ab/libvae21_xpf.so > "$OUT/${TAG}_fp32_ab.json" 2>> "$OUT/${TAG}_bench.err"

# 4. ncu: the launch list of the bench command, then one full capture per kernel (one GPU, never a multi-rank command)
NCU="ncu --clock-control none"
T=1200 run $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file "$OUT/${TAG}_bench_launches.csv" \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > "$OUT/${TAG}_ncu_bench.log" 2>&1
for p in fp16e4m3 bf16x3; do
  run python tools/run_once.py 1000000 $p 3 > "$OUT/${TAG}_run_once_$p.log" 2>&1 && \
  T=1200 run $NCU --set full --import-source on -k regex:vae21_tc -s 2 -c 1 -f -o "$OUT/${TAG}_tc_$p" \
    python tools/run_once.py 1000000 $p 3 > "$OUT/${TAG}_ncu_$p.log" 2>&1
done
run python tools/run_once.py 1000000 fp32 1 > "$OUT/${TAG}_run_once_fp32.log" 2>&1 && \
T=1200 run $NCU --set full --import-source on -k regex:pipe -c 1 -f -o "$OUT/${TAG}_fp32_pipe" \
  python tools/run_once.py 1000000 fp32 1 > "$OUT/${TAG}_ncu_fp32.log" 2>&1
echo "done: $(ls "$OUT" | grep -c "${TAG}") files under $OUT/" >&2
# back on the CPU box:  python tools/ncu_summary.py gpurun_out/<tag>_tc_fp16e4m3.ncu-rep profiles/<tag>_tc_fp16e4m3_1M_ncu_summary.md
#                       python tools/ncu_roles.py   gpurun_out/<tag>_tc_fp16e4m3.ncu-rep > profiles/<tag>_tc_fp16e4m3_roles.txt
