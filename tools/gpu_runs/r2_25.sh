mkdir -p gpurun_out
N=$1
nvidia-smi topo -m > gpurun_out/r2_25_topo.txt 2>&1
(lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c; nproc) > gpurun_out/r2_25_cpu.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29655 tools/hostlink_probe.py > gpurun_out/r2_25_hostlink.json 2> gpurun_out/r2_25_hostlink.err
cat gpurun_out/r2_25_hostlink.json; tail -2 gpurun_out/r2_25_hostlink.err; head -14 gpurun_out/r2_25_topo.txt | cut -c1-150; cat gpurun_out/r2_25_cpu.txt
