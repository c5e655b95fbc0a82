#!/bin/bash
# host topology of the GPU box: NUMA nodes, the CPUs this process may use, GPU <-> NUMA affinity (diagnostic)
echo "== nproc / affinity"; nproc; taskset -p $$ 2>/dev/null; grep -i "allowed" /proc/self/status
echo "== NUMA nodes"; for n in /sys/devices/system/node/node*; do echo "$n: cpus $(cat $n/cpulist) mem $(grep MemTotal $n/meminfo | awk '{print $4,$5}')"; done
echo "== GPUs"; nvidia-smi --query-gpu=index,pci.bus_id,name --format=csv,noheader
for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//'); do echo "$d numa_node $(cat /sys/bus/pci/devices/$d/numa_node 2>/dev/null)"; done
echo "== topo"; nvidia-smi topo -m 2>/dev/null | head -30
echo "== lscpu"; lscpu | grep -i "model name\|socket\|numa\|^cpu(s)"
echo "== mem"; free -g | head -2
