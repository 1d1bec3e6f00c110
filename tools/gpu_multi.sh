#!/bin/bash
# multi-GPU run (gpurun --gpus N -- tools/gpu_multi.sh N [tag]): the driver-style bench line at N ranks
# (frame-parallel headline + strips + cfg5 batch), the host-link ceiling at N ranks, topology.
N=${1:-2}; tag=${2:-r02}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 1200 bash -c "$(declare -f run); N=$N; run 29511 bench.py --gpus $N --steps 20 --warmup 3" > gpurun_out/bench_n${N}_$tag.json 2> gpurun_out/bench_n${N}_$tag.err
echo "bench exit $?"; tail -c 2500 gpurun_out/bench_n${N}_$tag.json; echo; grep -v "^\[W\|^W0\|^\*\*\*\|^$" gpurun_out/bench_n${N}_$tag.err | tail -8
timeout 300 bash -c "$(declare -f run); N=$N; run 29512 tools/pcie_ceiling.py" > gpurun_out/pcie_n${N}_$tag.log 2>&1; grep '^{' gpurun_out/pcie_n${N}_$tag.log
(nvidia-smi topo -m; nproc; free -g | head -2) > gpurun_out/topology_n${N}_$tag.txt 2>&1
