#!/bin/bash
# The multi-GPU measurements of one box size N: bench.py config 2 and 3 (as the driver launches them)
# and the config-5 all-pairs job.  Usage: bash tools/multi_gpu_suite.sh N [cli]
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_c2_n$N.json 2> gpurun_out/r02_bench_c2_n$N.err
$TR bench.py --gpus $N --config 3 --steps 3 --warmup 3 > gpurun_out/r02_bench_c3_n$N.json 2> gpurun_out/r02_bench_c3_n$N.err
$TR tools/config5_run.py --genomes 1000 --out gpurun_out/r02_config5_n$N.json > /dev/null 2> gpurun_out/r02_config5_n$N.err
if [ "$2" = "cli" ]; then
  $TR tools/config5_cli.py --genomes 1000 --out gpurun_out/r02_config5_cli_n$N.json > /dev/null 2> gpurun_out/r02_config5_cli_n$N.err
  python tools/config3_cli.py --gpus $N --genomes 8 --bases 3.1e9 --cpu-sample-bytes 256e6 --also-torchrun --out gpurun_out/r02_config3_cli_${N}gpu.json > /dev/null 2> gpurun_out/cfg3_${N}gpu.err
fi
python - <<PY
import json
for name in ("bench_c2", "bench_c3"):
    try:
        d = json.load(open("gpurun_out/r02_%s_n$N.json" % name))
        print(name, "N=$N value", round(d["value"], 3), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 3), round(d["e2e"]["ms_per_step"], 3), "numa", d["e2e"].get("host_numa_node"), "launches", d["gpu_launches"])
    except Exception as e:
        print(name, "FAILED", e)
try:
    d = json.load(open("gpurun_out/r02_config5_n$N.json"))
    print("config5 N=$N", {k: round(d[k], 4) for k in ("sketch_s", "gather_s", "to_planes_s", "pairs_s", "total_s")}, d["oracle_max_rel_err"], d["oracle_registers_equal"])
except Exception as e:
    print("config5 FAILED", e)
try:
    d = json.load(open("gpurun_out/r02_config5_cli_n$N.json"))
    print("config5 from files N=$N wall", round(d["allpairs_wall_s"], 2), {k: round(v, 3) for k, v in d["stages_rank0"].items()}, d["oracle_max_rel_err"])
except Exception as e:
    print("config5 cli: not run or failed", e)
try:
    d = json.load(open("gpurun_out/r02_config3_cli_${N}gpu.json"))
    print("cfg3 cli", {k: d.get(k) for k in ("tree_wall_s", "tree_wall_torchrun_s", "torchrun_cards_identical", "tree_cached_rerun_wall_s", "progressive_wall_s", "tree_speedup_vs_cpu_proxy")})
    print(json.dumps(d["tree_stages"]))
except Exception as e:
    print("cfg3 cli: not run or failed", e)
PY
