#!/usr/bin/env python
"""Benchmark of the learner hot path: GAE + ppo_error (forward AND backward) on a (T, B) trajectory batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Metric (BASELINE.json): learner transitions/sec, config D = Atari-PPO shape T=128, B=4096 per GPU, N=6 actions
(``pong_ppo_config.py``: gamma 0.99, lambda 0.95, clip 0.2, value clip on).  One step = one pass of
gae -> ppo_error forward -> backward(policy + 0.5 value - 0.01 entropy) over one batch of 524 288 transitions per GPU.

  value     inputs resident in HBM; the step's kernels (default: the one-launch gae+ppo_error step of csrc/colws.cu, its
            finalize_sums and the backward's verification launch; --three / --unfused: the separate operators) replayed
            as a CUDA graph; input/output buffer sets are rotated so that consecutive steps never find their data in the
            126 MB L2 (4 sets x 67 MB); timed with CUDA events on the launching stream between barrier + synchronize,
            max over ranks.
  e2e       the same step through the public API (di_engine_b200.gae_ppo_error / backward; --three: gae, ppo_error)
            starting from PINNED HOST buffers: per step H2D copy of every input, the kernels, D2H read of the loss.
  roofline  per-kernel CUDA-event timing of the dominant kernel against MEASURED_PEAKS.json (HBM copy bandwidth).
  cpu_baseline / --impl reference
            the reference algorithm on the host cores: oracle/rl_oracle.py, the torch-CPU restatement that is pinned
            bit-exact to ding.rl_utils (the reference is pure Python and cannot travel to the GPU box).

Multi-GPU (torchrun, one rank per GPU): the batch shards along B with no data-path exchange; the only collective is
one NCCL all-reduce per step of the packed loss scalars (mean-of-rank-means, as DI-engine's DDP does,
ding/utils/pytorch_ddp_dist_helper.py:38-47), issued on a side stream so it overlaps the next step's kernels.
Weak scaling: every rank holds a full T=128 x B=4096 shard.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_LEN, B_COLS, N_ACT = 128, 4096, 6
GAMMA, LAMBDA, CLIP = 0.99, 0.95, 0.2
W_VALUE, W_ENTROPY = 0.5, -0.01
ALG_BYTES_PER_TR = {'gae_ppo_fwd_grad': 128, 'gae': 24, 'ppo_fwd': 76, 'ppo_bwd': 100, 'ppo_fwd_grad': 104, 'ppo_bwd_check': 0, 'step': 128}
METRIC = 'learner transitions/sec (GAE+ppo_error fwd+bwd, T=128 x B=4096 per GPU)'


# ----------------------------------------------------------------------------------------------------------------
# synthetic batch
# ----------------------------------------------------------------------------------------------------------------
def make_batch(seed, T=T_LEN, B=B_COLS, N=N_ACT):
    g = torch.Generator().manual_seed(seed)
    value = torch.randn(T, B, generator=g)
    done = (torch.rand(T, B, generator=g) < 0.01).float()
    next_value = torch.cat([value[1:], torch.randn(1, B, generator=g)], 0)
    next_value = torch.where(done.bool(), torch.randn(T, B, generator=g), next_value).contiguous()
    reward = torch.randn(T, B, generator=g)
    traj = done.clone()
    traj[-1] = 1.0
    logit_new = torch.randn(T * B, N, generator=g)
    logit_old = logit_new + 0.1 * torch.rand(T * B, N, generator=g)
    action = torch.randint(0, N, (T * B, ), generator=g)
    value_new = torch.randn(T * B, generator=g)
    value_old = value_new + 0.1 * torch.rand(T * B, generator=g)
    return_ = torch.randn(T * B, generator=g)
    return dict(value=value, next_value=next_value, reward=reward, done=done, traj_flag=traj, logit_new=logit_new,
                logit_old=logit_old, action=action, value_new=value_new, value_old=value_old, return_=return_)


def batch_bytes(b):
    return sum(v.numel() * v.element_size() for v in b.values())


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_step(orc, b):
    adv = orc.gae(b['value'], b['next_value'].clone(), b['reward'], b['done'], b['traj_flag'], GAMMA, LAMBDA)
    ln = b['logit_new'].detach().requires_grad_(True)
    vn = b['value_new'].detach().requires_grad_(True)
    p, v, e, k, akl, cf = orc.ppo_error(ln, b['logit_old'], b['action'], vn, b['value_old'], adv.reshape(-1),
                                        b['return_'], None, None, CLIP, True, None)
    (p + W_VALUE * v + W_ENTROPY * e).backward()
    return float(p.detach())


def usable_cores():
    """Host cores this process may actually use: affinity mask, clipped by the cgroup CPU quota if there is one."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()[:2]
        if quota != 'max':
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except (OSError, ValueError):
        pass
    return n


def pick_threads(orc, b):
    """All the host threads the reference can USE: the fastest of {usable, 64, 32, 16, 8} torch intra-op threads
    (over-subscribing a throttled container makes torch slower, not faster)."""
    usable = usable_cores()
    best, best_t = None, None
    for n in sorted({usable, 64, 32, 16, 8}):
        if n > usable:
            continue
        torch.set_num_threads(n)
        cpu_step(orc, b)
        t0 = time.perf_counter()
        cpu_step(orc, b)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
        if dt > 4 * best_t:
            break
    return best


def run_cpu(steps, warmup):
    from oracle import rl_oracle
    b = make_batch(0)
    cores = pick_threads(rl_oracle, b)
    torch.set_num_threads(cores)
    for _ in range(warmup):
        cpu_step(rl_oracle, b)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_step(rl_oracle, b)
        times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return dict(value=T_LEN * B_COLS / med, ms_per_step=med * 1e3, total_s=sum(times), cores=cores,
                threads=torch.get_num_threads())


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.samples = []
        self.proc = None
        self.idx = gpu_index
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for s in self.samples:
            f = [x.strip() for x in s.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except (KeyError, ValueError):
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class DeviceStep:
    """One learner step on device-resident buffers, via the tensor-level layer under the public API."""

    def __init__(self, host_batch, dev, fused=True):
        from di_engine_b200 import ops
        self.ops = ops
        self.fused = fused
        self.hint = torch.tensor([1.0, W_VALUE, W_ENTROPY, 0.0], device=dev)
        self.g_used = torch.zeros(4, device=dev)
        self.b = {k: v.to(dev) for k, v in host_batch.items()}
        self.nv0 = self.b['next_value'].clone()
        self.S = T_LEN * B_COLS
        self.g_p = torch.tensor(1.0, device=dev)
        self.g_v = torch.tensor(W_VALUE, device=dev)
        self.g_e = torch.tensor(W_ENTROPY, device=dev)
        self.adv = torch.empty_like(self.b['value'])
        self.out = torch.zeros(8, device=dev)
        self.grad_logit = torch.empty_like(self.b['logit_new'])
        self.grad_value = torch.empty_like(self.b['value_new'])
        self.ws = ops.workspace(torch.device(dev))

    def gae(self):
        b, o = self.b, self.ops
        rc = o.lib().b200rl_gae(o.ptr(b['value']), o.ptr(b['next_value']), o.ptr(b['reward']), o.ptr(b['done']),
                                o.ptr(b['traj_flag']), o.ptr(self.adv), T_LEN, B_COLS, 1, GAMMA, LAMBDA, 1,
                                o.stream_ptr())
        assert rc == 0, rc

    def ppo_fwd(self):
        b, o = self.b, self.ops
        rc = o.lib().b200rl_ppo_fwd(o.ptr(b['logit_new']), o.ptr(b['logit_old']), None, o.ptr(b['action']),
                                    o.ptr(b['value_new']), o.ptr(b['value_old']), o.ptr(self.adv), o.ptr(b['return_']),
                                    None, self.S, 1, N_ACT, CLIP, 1, 0.0, 1, o.ptr(self.out), o.ptr(self.ws),
                                    self.ws.numel() * 4, o.stream_ptr())
        assert rc == 0, rc

    def ppo_bwd(self):
        b, o = self.b, self.ops
        rc = o.lib().b200rl_ppo_bwd(o.ptr(b['logit_new']), o.ptr(b['logit_old']), None, o.ptr(b['action']),
                                    o.ptr(b['value_new']), o.ptr(b['value_old']), o.ptr(self.adv), o.ptr(b['return_']),
                                    None, self.S, 1, N_ACT, CLIP, 1, 0.0, 1, o.ptr(self.g_p), o.ptr(self.g_v),
                                    o.ptr(self.g_e), None, None, None, o.ptr(self.grad_logit), o.ptr(self.grad_value),
                                    o.stream_ptr())
        assert rc == 0, rc

    def ppo_fwd_grad(self):
        b, o = self.b, self.ops
        rc = o.lib().b200rl_ppo_fwd_grad(o.ptr(b['logit_new']), o.ptr(b['logit_old']), None, o.ptr(b['action']),
                                         o.ptr(b['value_new']), o.ptr(b['value_old']), o.ptr(self.adv),
                                         o.ptr(b['return_']), None, self.S, 1, N_ACT, CLIP, 1, 0.0, 1, o.ptr(self.hint),
                                         o.ptr(self.g_used), o.ptr(self.out), o.ptr(self.grad_logit),
                                         o.ptr(self.grad_value), o.ptr(self.ws), self.ws.numel() * 4, o.stream_ptr())
        assert rc == 0, rc

    def ppo_bwd_check(self):
        b, o = self.b, self.ops
        rc = o.lib().b200rl_ppo_bwd(o.ptr(b['logit_new']), o.ptr(b['logit_old']), None, o.ptr(b['action']),
                                    o.ptr(b['value_new']), o.ptr(b['value_old']), o.ptr(self.adv), o.ptr(b['return_']),
                                    None, self.S, 1, N_ACT, CLIP, 1, 0.0, 1, o.ptr(self.g_p), o.ptr(self.g_v),
                                    o.ptr(self.g_e), None, o.ptr(self.g_used), o.ptr(self.hint),
                                    o.ptr(self.grad_logit), o.ptr(self.grad_value), o.stream_ptr())
        assert rc == 0, rc

    def gae_ppo_fwd_grad(self):
        b, o = self.b, self.ops
        rc = o.lib().b200rl_gae_ppo_fwd_grad(
            o.ptr(b['value']), o.ptr(b['next_value']), o.ptr(b['reward']), o.ptr(b['done']), o.ptr(b['traj_flag']),
            T_LEN, B_COLS, GAMMA, LAMBDA, 1, o.ptr(b['logit_new']), o.ptr(b['logit_old']), None, o.ptr(b['action']),
            o.ptr(b['value_new']), o.ptr(b['value_old']), o.ptr(b['return_']), None, N_ACT, CLIP, 1, 0.0, 1,
            o.ptr(self.hint), o.ptr(self.g_used), o.ptr(self.adv), o.ptr(self.out), o.ptr(self.grad_logit),
            o.ptr(self.grad_value), o.ptr(self.ws), self.ws.numel() * 4, o.stream_ptr())
        assert rc == 0, rc

    def kernels(self):
        if self.fused == 'onepass':
            return [('gae_ppo_fwd_grad', self.gae_ppo_fwd_grad), ('ppo_bwd_check', self.ppo_bwd_check)]
        if self.fused:
            return [('gae', self.gae), ('ppo_fwd_grad', self.ppo_fwd_grad), ('ppo_bwd_check', self.ppo_bwd_check)]
        return [('gae', self.gae), ('ppo_fwd', self.ppo_fwd), ('ppo_bwd', self.ppo_bwd)]

    def __call__(self):
        for _, k in self.kernels():
            k()


def run_gpu(args):
    import torch.distributed as dist
    import di_engine_b200 as b2

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        raise SystemExit('launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)' % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = 'cuda:%d' % local
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, args.warmup
    NSETS = 4
    mode = False if args.unfused else (True if args.three else 'onepass')
    sets = [DeviceStep(make_batch(1000 * rank + i), dev, fused=mode) for i in range(NSETS)]
    step_bytes = ALG_BYTES_PER_TR['step'] * T_LEN * B_COLS
    side = torch.cuda.Stream()
    main = torch.cuda.Stream()
    from di_engine_b200.parallel import LossAllReduce, P2PLossAllReduce
    reducers, exchange = None, 'none'
    if world > 1:
        if args.collective in ('auto', 'p2p'):
            try:  # one small kernel over NVLink peer memory (symmetric memory); falls back to NCCL if unavailable
                reducers = [P2PLossAllReduce(6, dev) for _ in range(NSETS)]
                exchange = 'p2p'
            except Exception as e:
                if args.collective == 'p2p':
                    raise
                if rank == 0:
                    print('bench: peer-memory all-reduce unavailable (%s); using NCCL' % e, file=sys.stderr)
        if reducers is None:
            reducers = [LossAllReduce(6, dev) for _ in range(NSETS)]
            exchange = 'nccl'

    def exchange_losses(j):
        """mean over ranks of set j's six loss scalars (mean of rank means), on the current stream"""
        if exchange == 'p2p':
            reducers[j].reduce(sets[j].out)
        else:
            reducers[j].buf.copy_(sets[j].out[:6], non_blocking=True)
            reducers[j].reduce()

    # ---- correctness guard: first set against the CPU oracle on rank 0 (outside every timed region) ----------------
    if rank == 0:
        from oracle import rl_oracle
        hb = make_batch(0)
        s0 = sets[0]
        s0()
        torch.cuda.synchronize()
        adv_ref = rl_oracle.gae(hb['value'], hb['next_value'].clone(), hb['reward'], hb['done'], hb['traj_flag'], GAMMA,
                                LAMBDA)
        assert torch.equal(s0.adv.cpu(), adv_ref), 'gae parity broken'
        s0.b['next_value'].copy_(s0.nv0)

    # ---- capture one graph per buffer set -------------------------------------------------------------------------
    # N > 1: the graph of step j also carries, on a forked branch, the all-reduce of the loss scalars of step j-1
    # (software-pipelined: the collective of one step overlaps the kernels of the next; one graph launch per step)
    graphs = []
    collective_mode = 'none'
    with torch.cuda.stream(main):
        for s in sets:
            s()
            s()
        if world > 1:
            for j in range(NSETS):
                exchange_losses(j)
        main.synchronize()

        def record_step(j, with_collective):
            if with_collective:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    exchange_losses((j - 1) % NSETS)
            sets[j]()
            if with_collective:
                main.wait_stream(side)

        def capture(with_collective):
            """graphs[j]: one step on buffer set j; graphs[NSETS]: the NSETS steps 0..NSETS-1 back to back (one host launch
            per NSETS steps -- the per-launch host cost of a ~25 us step is not negligible)"""
            gs = []
            for j in range(NSETS):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=main):
                    record_step(j, with_collective)
                gs.append(g)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main):
                for j in range(NSETS):
                    record_step(j, with_collective)
            gs.append(g)
            return gs

        if world > 1:
            try:
                graphs = capture(True)
                collective_mode = 'in-graph'
            except Exception as e:  # NCCL capture unavailable: keep the collective eager on the side stream
                if rank == 0:
                    print('bench: NCCL graph capture failed (%s); eager side-stream all-reduce' % e, file=sys.stderr)
                torch.cuda.synchronize()
                graphs = capture(False)
                collective_mode = 'eager'
        else:
            graphs = capture(False)
    torch.cuda.synchronize()

    def device_loop(n):
        """exactly n steps, buffer sets in the order 0,1,..,NSETS-1,0,1,.."""
        if collective_mode != 'eager':
            q, r = divmod(n, NSETS)
            for _ in range(q):
                graphs[NSETS].replay()
            for j in range(r):
                graphs[j].replay()
        for i in range(n if collective_mode == 'eager' else 0):
            j = i % NSETS
            graphs[j].replay()
            if collective_mode == 'eager':
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    exchange_losses(j)
        if collective_mode == 'in-graph':  # the last step's scalars (every earlier one rode in the next step's graph)
            exchange_losses((n - 1) % NSETS)
        elif collective_mode == 'eager':
            main.wait_stream(side)

    with torch.cuda.stream(main):
        device_loop(max(W, 3))
        # pre-heat: ~0.25 s of the same steps (untimed) so SM/memory clocks and caches are in steady state -- one step is
        # only ~30 us, far shorter than the clock governor's reaction time.  A FIXED step count: with a collective in the
        # graph every rank must launch exactly the same number of steps.
        for _ in range(8):
            device_loop(1000)
            torch.cuda.synchronize()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(main)
        device_loop(K)
        e1.record(main)
        barrier()
        dev_ms = e0.elapsed_time(e1)

        # ---- the same step launched eagerly (one ctypes call per launch, no graph): host-bound, reported next to the replay
        eager_ms = None
        try:
            n_eager = 200
            for i in range(8):
                sets[i % NSETS]()
            main.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(main)
            for i in range(n_eager):
                sets[i % NSETS]()
            g1.record(main)
            main.synchronize()
            eager_ms = g0.elapsed_time(g1) / n_eager
        except Exception as e:  # never let the secondary figure break the benchmark
            if rank == 0:
                print('bench: eager-launch timing skipped (%s)' % e, file=sys.stderr)

        # ---- per-kernel timing: each kernel alone, back to back over the rotated buffer sets, replayed as a graph so
        # that launch gaps of the host do not enter the figure (CUDA events on the launching stream)
        names = [n for n, _ in sets[0].kernels()]
        per = {}
        reps = max(100, min(K, 2000) // NSETS)
        for ki, name in enumerate(names):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main):
                for s in sets:
                    s.kernels()[ki][1]()
            for _ in range(3):
                g.replay()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            main.synchronize()
            k0.record(main)
            for _ in range(reps):
                g.replay()
            k1.record(main)
            main.synchronize()
            per[name] = [k0.elapsed_time(k1) / (reps * NSETS)]
        clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end through the public API from pinned host buffers ------------------------------------------------
    packed = not args.e2e_separate_copies
    if packed:  # one pinned buffer + one device buffer per slot: the H2D transfer of a step is a single copy
        slots = [b2.PackedBatch(make_batch(2000 * rank + i), dev) for i in range(2)]
        host = slots
        h2d = slots[0].payload_bytes()
    else:
        host = [{k: v.pin_memory() for k, v in make_batch(2000 * rank + i).items()} for i in range(2)]
        h2d = batch_bytes(host[0])

    # The host side is a two-deep prefetching loader (what DI-engine's CudaFetcher, ding/torch_utils/data_helper.py:543,
    # does for the learner): the H2D copy of step i+1 is enqueued on a copy stream before step i's result is read back,
    # so PCIe transfer and kernels overlap.  Every byte of every step is still copied inside the timed region.
    copy_stream = torch.cuda.Stream()

    def upload(hb):
        if packed:
            return hb.upload(copy_stream)
        with torch.cuda.stream(copy_stream):
            d = {k: v.to(dev, non_blocking=True) for k, v in hb.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return d, ev

    def e2e_compute(d, ev):
        torch.cuda.current_stream().wait_event(ev)
        if not packed:
            for v in d.values():
                v.record_stream(torch.cuda.current_stream())
        ln = d['logit_new'].requires_grad_(True)
        vn = d['value_new'].requires_grad_(True)
        gd = b2.gae_data(d['value'], d['next_value'], d['reward'], d['done'], d['traj_flag'])
        if not (args.three or args.unfused):
            adv, loss, info = b2.gae_ppo_error(
                gd, b2.ppo_data(ln, d['logit_old'], d['action'], vn, d['value_old'], None, d['return_'], None, None),
                GAMMA, LAMBDA, CLIP, True, None)
        else:
            adv = b2.gae(gd, GAMMA, LAMBDA)
            loss, info = b2.ppo_error(
                b2.ppo_data(ln, d['logit_old'], d['action'], vn, d['value_old'], adv.view(-1), d['return_'], None,
                            None), CLIP, True, None)
        total = loss.policy_loss + W_VALUE * loss.value_loss + W_ENTROPY * loss.entropy_loss
        total.backward()
        return total

    def e2e_loop(n):
        nxt = upload(host[0])
        last = None
        for i in range(n):
            cur = nxt
            if i + 1 < n:
                nxt = upload(host[(i + 1) % 2])
            total = e2e_compute(*cur)
            last = total.item()  # D2H read of the step's result (ppo_info already cost one 8-byte read)
        return last

    e2e_steps = max(5, min(K, 20))
    e2e_loop(3)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(e2e_steps)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)

    # ---- max over ranks --------------------------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()

    if rank == 0:
        peak, peak_src = load_peaks()
        tr_per_step = T_LEN * B_COLS * world
        ms_step = dev_ms / K
        value = tr_per_step / (ms_step * 1e-3)
        kmean = {n: statistics.mean(v) for n, v in per.items()}
        dom = max(kmean, key=kmean.get)
        dom_bytes = ALG_BYTES_PER_TR[dom] * T_LEN * B_COLS
        achieved = dom_bytes / (kmean[dom] * 1e-3) / 1e9
        step_achieved = step_bytes / (ms_step * 1e-3) / 1e9
        traffic = None
        try:  # DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/)
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'))).get(dom)
            if tr:
                traffic = tr['dram_read'] + tr['dram_write']
        except (OSError, ValueError, KeyError):
            pass
        cpu = run_cpu(steps=8, warmup=2) if world == 1 else None
        e2e_value = tr_per_step / (e2e_ms / e2e_steps * 1e-3)
        line = {
            'metric': METRIC, 'value': value, 'unit': 'transitions/s', 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {
                'workload': 'configs[3] Atari PPO gae+ppo_error T=128 B=4096 N=6 per GPU (B sharded across GPUs; obs '
                            '[4,84,84] is not read by any kernel on this path and is not materialised)',
                'transitions_per_step_per_gpu': T_LEN * B_COLS, 'gamma': GAMMA, 'lambda': LAMBDA, 'clip_ratio': CLIP,
                'loss_mix': [1.0, W_VALUE, W_ENTROPY], 'parallelism': 'dp%d' % world,
                'l2_policy': 'inputs rotated over %d buffer sets of 67 MB (> 126 MB L2) between consecutive steps' %
                             NSETS,
                'launch': 'CUDA graph replay, %d steps per graph launch, %d kernels per step (%s)' % (NSETS, len(names), ', '.join(names)),
                'collective': 'none' if world == 1 else ('one all-reduce (mean) of the 6 loss scalars per step, %s, %s, overlapping the next step' % ('NVLink peer-memory kernel b200rl_p2p_allreduce_mean' if exchange == 'p2p' else 'NCCL', collective_mode)),
            },
            'roofline': {
                'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                'alg_bytes_per_launch': dom_bytes, 'kernel_ms': kmean,
                'step': {'alg_bytes': step_bytes, 'achieved': step_achieved, 'frac': step_achieved / peak},
            },
            'cpu_baseline': None if cpu is None else {
                'value': cpu['value'], 'unit': 'transitions/s', 'cores': cpu['cores'], 'kind': 'port',
                'sample': 'full T=128 x B=4096 batch, median of 8 steps after 2 warm-up (%.1f s CPU), %s' %
                          (cpu['total_s'], cpu_model()),
                'ms_per_step': cpu['ms_per_step'],
            },
            'e2e': {'value': e2e_value, 'unit': 'transitions/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 12,
                    'ms_per_step': e2e_ms / e2e_steps, 'steps': e2e_steps},
            'eager_ms_per_step': eager_ms,  # same launches without the CUDA graph (host-bound; `value` is the graph replay)
            'gpu_launches': (len(names) + 1) * K,  # + the finalize_sums launch behind every loss-reducing kernel
            'clocks': clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # graphs that captured NCCL work must be gone before the communicator is torn down; then leave without waiting on
        # NCCL's own teardown (a destroy_process_group after captured collectives has been seen to hang)
        del graphs
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = run_cpu(steps=args.steps, warmup=max(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': r['value'], 'unit': 'transitions/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': max(args.warmup, 1), 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'configs[3] Atari PPO gae+ppo_error T=128 B=4096 N=6 (one batch, host cores)',
                   'parallelism': 'cpu'},
        'cpu_baseline': {'value': r['value'], 'unit': 'transitions/s', 'cores': r['cores'], 'kind': 'port',
                         'sample': 'full T=128 x B=4096 batch per step, %d torch threads, %s' %
                                   (r['threads'], cpu_model())},
        'e2e': {'value': r['value'], 'unit': 'transitions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--collective', default='auto', choices=['auto', 'p2p', 'nccl'],
                    help='N>1: exchange of the loss scalars (auto = NVLink peer-memory kernel, NCCL if unavailable)')
    ap.add_argument('--unfused', action='store_true', help='separate gae / ppo forward / ppo backward kernels')
    ap.add_argument('--three', action='store_true',
                    help='gae, fused ppo forward+grad, verification as three kernels (default: the one-launch gae+ppo '
                         'step of csrc/colws.cu + verification)')
    ap.add_argument('--onepass', action='store_true', help='accepted for compatibility: the one-launch step is the default')
    ap.add_argument('--e2e-separate-copies', action='store_true',
                    help='e2e: one pinned tensor and one H2D copy per input (default: di_engine_b200.PackedBatch, one copy)')
    args = ap.parse_args()
    if args.impl == 'reference':
        if args.steps > 400:
            args.steps = 400  # bounded: ~50 ms of host work per step
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
