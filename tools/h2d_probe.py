"""Diagnostic: pinned host -> device copy bandwidth as a function of the CPU / NUMA placement of the pinned allocation.

    python tools/h2d_probe.py [gpu_index]

Prints the topology the driver reports (nvidia-smi topo -m, NVML cpu affinity of the GPU, the NUMA nodes of the host) and
the H2D bandwidth of a 46 MB and a 512 MB copy with the process pinned to each NUMA node in turn."""
import os
import subprocess
import sys
import time

import torch


def numa_nodes():
    out = {}
    base = '/sys/devices/system/node'
    if not os.path.isdir(base):
        return out
    for d in sorted(os.listdir(base)):
        if d.startswith('node') and d[4:].isdigit():
            try:
                cl = open(os.path.join(base, d, 'cpulist')).read().strip()
            except OSError:
                continue
            cpus = set()
            for part in cl.split(','):
                if '-' in part:
                    a, b = part.split('-')
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            out[int(d[4:])] = cpus
    return out


def bw(nbytes, dev, reps=20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    dev = 'cuda:%d' % gpu
    torch.cuda.set_device(gpu)
    try:
        print(subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True).stdout)
    except OSError:
        pass
    nodes = numa_nodes()
    print('numa nodes:', {k: '%d cpus' % len(v) for k, v in nodes.items()})
    allowed = os.sched_getaffinity(0)
    print('affinity of this process: %d cpus' % len(allowed))
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [i * 64 + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        print('NVML cpu affinity of GPU %d: %d cpus, first %s' % (gpu, len(cpus), cpus[:4]))
        try:
            print('NVML numa node of GPU:', pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception as e:
            print('nvmlDeviceGetNumaNodeId:', e)
    except Exception as e:
        print('pynvml unavailable:', e)
    print('default placement: 46MB %.1f GB/s, 512MB %.1f GB/s' % (bw(46 << 20, dev), bw(512 << 20, dev)))
    for n, cpus in nodes.items():
        use = cpus & allowed
        if not use:
            continue
        os.sched_setaffinity(0, use)
        time.sleep(0.05)
        print('pinned to node %d (%d cpus): 46MB %.1f GB/s, 512MB %.1f GB/s' % (n, len(use), bw(46 << 20, dev),
                                                                                 bw(512 << 20, dev)))
    os.sched_setaffinity(0, allowed)


if __name__ == '__main__':
    main()
