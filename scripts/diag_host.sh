#!/bin/bash
# host topology facts of the GPU box (for the e2e scaling notes): NUMA nodes, cores, PCIe placement of the GPUs
nproc; ls /sys/devices/system/node/ 2>/dev/null | head; cat /sys/devices/system/node/node*/cpulist 2>/dev/null
nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class; then echo $d $(cat $d/numa_node); fi; done
nvidia-smi topo -m 2>/dev/null | head -20
python -c "import os; print('affinity', len(os.sched_getaffinity(0)))"
free -g | head -2
