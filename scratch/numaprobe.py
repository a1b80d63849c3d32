import sys, torch, os
sys.path.insert(0, '/root/repo')
import bench
print('nodes', {k: len(v) for k, v in bench.numa_cpulists().items()})
print('sysfs/probe ->', bench.pin_to_gpu_numa_node(torch, 0))
os.system('nvidia-smi topo -m | head -8')
