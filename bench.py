#!/usr/bin/env python
"""bench.py -- ConvBPDN ADMM iterations/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one ADMM iteration (sporco/admm/admm.py:331-377) of ConvBPDN on the metric
configuration: 256x256 images, M=64 filters of 8x8, K=32 images per GPU, float32, lambda=0.1,
AutoRho enabled (class default), FastSolve=True (residuals + rho update every iteration, no
objective -- SURVEY.md section 8d "mode B").  Inputs are synthetic (seeded normal).

Own arm: `value` = K iterations / device time (CUDA events on the library's stream, state
resident in HBM, max over ranks); `e2e` = the same through the public class with host
arrays in and the coefficient maps back out; `roofline` = the dominant kernel's algorithmic
bytes / its event-timed duration against MEASURED_PEAKS.json; `cpu_baseline` = the numpy
oracle (a restatement of the reference, bit-identical to it here) on the full 32-image batch for a
few iterations; `configs` = short runs of BASELINE.json's other configurations plus `modes` (the metric
configuration in the other two timing modes of SURVEY.md section 8d: class defaults with statistics every iteration, and
FastSolve without AutoRho); at N > 1
`parity_check` = a small sharded solve against the oracle on the whole batch (the run fails if
it disagrees).  Reference arm (`--impl reference`): the oracle on the host cores, full batch, the
same warm-up / timed iterations of one solve as the device arm.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N0 = N1 = 256
M = 64
K_PER_GPU = 32
HD = 8
LMBDA = 0.1
METRIC = 'ConvBPDN ADMM iterations/sec (256x256, M=64 filters, batch=32 images per GPU, float32)'
OPT_BENCH = {'RelStopTol': 0.0, 'FastSolve': True, 'AutoRho': {'Enabled': True}}
WORKLOAD = ('admm.cbpdn.ConvBPDN 256x256, 8x8x64 dict, 32 images per GPU, lambda=0.1, AutoRho on (period 1), '
            'RelaxParam 1.8, FastSolve (mode B)')


def make_inputs(k_images, seed=12345):
    rng = np.random.default_rng(seed)
    D = rng.standard_normal((HD, HD, M)).astype(np.float32)
    D /= np.sqrt(np.sum(D ** 2, axis=(0, 1), keepdims=True))
    S = rng.standard_normal((N0, N1, k_images)).astype(np.float32)
    return D, S


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.FIELDS,
                 '--format=csv,noheader,nounits', '-lms', '50'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                  'sw_power_cap'), f[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_solve(k_images, warm, steps, kind, workers):
    """Time the oracle (oracle/cbpdn_oracle.py: numpy restatement of the reference, pinned bit for bit to
    it) on `k_images` images of the metric workload: ONE solve of warm + steps iterations with a time stamp
    after every iteration; returns seconds per iteration over the last `steps` of them -- the same
    iterations of the same solve that the device arm times."""
    from oracle import cbpdn_oracle as orc
    D, S = make_inputs(k_images)
    tm = {}
    opt = dict(OPT_BENCH)
    opt['MaxMainIter'] = warm + steps
    orc.admm_convbpdn(D, S, LMBDA, opt=opt, dimK=1, fft=orc.FFTBackend(kind, workers), timing=tm)
    te = tm['iter_end']
    assert len(te) == warm + steps
    t0 = te[warm - 1] if warm > 0 else 0.0
    return (te[-1] - t0) / steps


def run_reference(args):
    """The reference's CPU implementation of the path (the oracle port; sporco itself is pure Python
    over numpy / pyfftw and cannot travel to the GPU box) on the host cores, FULL metric workload:
    32 images, `warmup` untimed + `steps` timed iterations of one solve."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # whole job at N GPUs = N batches of 32 images (weak scaling); the CPU arm runs them as one
    # ConvBPDN object of 32 N images, capped at 64 images (N <= 2 exact; beyond, the 64-image time per
    # batch is used for the remaining batches and the line says so)
    n_batches = max(1, args.gpus)
    run_batches = min(n_batches, 2)
    t_wall = time.perf_counter()
    per_iter = cpu_solve(K_PER_GPU * run_batches, args.warmup, args.steps, 'scipy', cores)
    t_wall = time.perf_counter() - t_wall
    value = run_batches / per_iter          # batch-iterations per second, as the device arm counts them
    per_iter = per_iter / run_batches * n_batches
    sample = ('oracle (numpy restatement of sporco.admm.cbpdn.ConvBPDN, bit-identical to the reference in '
              'the build container) with scipy.fft workers=%d standing in for the reference\'s multi-threaded '
              'pyfftw; %d images (%s), %d warm-up + %d timed iterations of one solve (%.1f s wall clock in all)'
              % (cores, K_PER_GPU * run_batches,
                 'the full job' if run_batches == n_batches else
                 'of the %d of the %d-GPU job; per-batch rate assumed the same for the rest' % (K_PER_GPU * n_batches, n_batches),
                 args.warmup, args.steps, t_wall))
    out = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'iterations/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1000.0 * per_iter, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample_images': K_PER_GPU * run_batches,
                   'images_total': K_PER_GPU * n_batches, 'same_config': run_batches == n_batches},
        'cpu_baseline': {'value': value, 'unit': 'iterations/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'iterations/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
    }
    print(json.dumps(out))


def multi_gpu_parity(dist, rank, world, device):
    """Untimed check run by `bench.py --gpus N>1`: a small problem sharded over the ranks (2 images each)
    against the oracle on the WHOLE batch -- the coefficient maps of this rank's images, the rho
    trajectory and the iteration at which the (global) stopping test fires must agree
    (sporco/admm/admm.py:462-486: the norms are global over all images)."""
    import torch
    from sporco_b200.admm import cbpdn
    from oracle import cbpdn_oracle as orc
    rng = np.random.default_rng(77)
    Dp = rng.standard_normal((6, 6, 8)).astype(np.float32)
    Sp = rng.standard_normal((64, 64, 2 * world)).astype(np.float32)
    opt = {'MaxMainIter': 60, 'RelStopTol': 4e-3, 'AutoRho': {'Enabled': True}}
    b = cbpdn.ConvBPDN(Dp, np.ascontiguousarray(Sp[:, :, 2 * rank:2 * rank + 2]), 0.05,
                       cbpdn.ConvBPDN.Options(opt), dimK=1, device=device)
    b.attach_process_group(dist)
    Y = b.solve()
    its = b.getitstat()
    # The check is against the float64 oracle on the same (float32) inputs: over these 60 AutoRho iterations the
    # reference's own float32 run drifts 3e-4 from its float64 run, so agreement with the float32 oracle to 1e-4
    # is not a property any float32 implementation has; the rho trajectory and the stop iteration are compared
    # with the float32 oracle (they are decided in float32).
    r = orc.admm_convbpdn(Dp, Sp, 0.05, opt=opt, dimK=1)
    opt64 = dict(opt, MaxMainIter=len(r.itstat), RelStopTol=0.0)        # the same number of iterations
    r64 = orc.admm_convbpdn(Dp.astype(np.float64), Sp.astype(np.float64), 0.05, opt=opt64, dimK=1)
    sl = (slice(None), slice(None), slice(None), slice(2 * rank, 2 * rank + 2), slice(None))
    Yr = r.Y[sl].reshape(Y.shape)
    Y64 = r64.Y[sl].reshape(Y.shape)
    rel_y = float(np.linalg.norm((Y - Y64).ravel()) / max(np.linalg.norm(Y64.ravel()), 1e-30))
    rel_y32 = float(np.linalg.norm((Y - Yr).ravel()) / max(np.linalg.norm(Yr.ravel()), 1e-30))
    drift = float(np.linalg.norm((r.Y - r64.Y).ravel()) / max(np.linalg.norm(r64.Y.ravel()), 1e-30))
    rho_ref = np.array([row[8] for row in r.itstat], dtype=np.float64)
    rho_own = np.asarray(its.Rho, dtype=np.float64)
    same_n = len(rho_own) == len(rho_ref)
    rho_rel = float(np.max(np.abs(rho_own - rho_ref) / rho_ref)) if same_n else float('inf')
    t = torch.tensor([rel_y, rho_rel, 0.0 if same_n else 1.0, rel_y32], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rel_y, rho_rel, bad_n, rel_y32 = (float(x) for x in t.tolist())
    ok = bool(rel_y < 1e-4 and rho_rel < 1e-4 and bad_n == 0.0)
    del b
    return {'n': world, 'problem': '64x64, 6x6x8 dictionary, %d images (2 per rank), AutoRho, stop at RelStopTol 4e-3'
            % (2 * world), 'iterations': int(len(rho_ref)), 'stop_iteration_equal': bad_n == 0.0,
            'rel_Y': rel_y, 'rel_Y_reference': 'float64 oracle on the same inputs',
            'rel_Y_vs_float32_oracle': rel_y32, 'oracle_float32_vs_float64': drift,
            'rho_rel': rho_rel, 'tol': 1e-4, 'ok': ok}


def recorded_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r02_traffic.json,
    written by tools/ncu_summary.py from `ncu --set full`: dram__bytes_read.sum + dram__bytes_write.sum)."""
    p = os.path.join(ROOT, 'profiles', 'r02_traffic.json')
    try:
        rec = json.load(open(p))
    except Exception:
        return None, None
    for name, v in rec.get('kernels', {}).items():
        if name.startswith(kernel):
            return float(v['dram_bytes']), 'profiles/r02_traffic.json (%s; %s)' % (name, rec.get('source', ''))
    return None, None


def run_b200(args):
    import torch
    from sporco_b200 import _lib
    from sporco_b200.admm import cbpdn

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit('launch with torch.distributed.run for --gpus > 1')
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    torch.cuda.set_device(local_rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    D, S = make_inputs(K_PER_GPU, seed=12345 + rank)
    opt = dict(OPT_BENCH)
    opt['MaxMainIter'] = args.steps
    b = cbpdn.ConvBPDN(D, S, LMBDA, cbpdn.ConvBPDN.Options(opt), dimK=1, device=local_rank)
    if world > 1:
        b.attach_process_group(dist)
    h = b._h
    h.admm_configure(**b._admm_config())

    # ---- device-resident throughput
    if args.warmup > 0:
        h.admm_iterate(args.warmup, False)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    _, done, _ = h.admm_iterate(args.steps, False)
    barrier()
    ms, launches = h.admm_last_timing()
    clocks = sampler.stop() if rank == 0 else None
    assert done == args.steps, 'solver stopped early (%d of %d)' % (done, args.steps)
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    # whole-job aggregate: every rank advances its own 32-image batch by one iteration per
    # step, so N GPUs complete N batch-iterations per step (weak scaling)
    value = world * args.steps / (ms_max / 1000.0)
    img_iter_per_s = world * K_PER_GPU * args.steps / (ms_max / 1000.0)

    # ---- per-kernel timing for the roofline (event between kernels; same stream)
    prof_n = min(20, args.steps)
    kms = h.admm_profile(prof_n)
    kms = [x / prof_n for x in kms]
    b_r = 4.0 * N0 * N1 * K_PER_GPU * M
    zt = 8.0 * N0 * (N1 // 2 + 1) * K_PER_GPU * M
    small = 8.0 * N0 * (N1 // 2 + 1) * (M + K_PER_GPU)
    sched = h.admm_schedule_info()
    # algorithmic bytes per launch.  With cross-iteration fusion the prox
    # kernel also writes the next x-step's row spectra and the row-forward launch only has work
    # in the iterations where rho changed (none in the steady state that is timed here).
    alg_bytes = {'k_row_fwd': (2 * b_r + zt) if not sched['fused'] else 0.0,
                 'k_col': 2 * zt + small,
                 'k_row_inv_prox': zt + 4 * b_r + (zt if sched['fused'] else 0.0)}
    names = ['k_row_fwd', 'k_col', 'k_row_inv_prox', 'k_admm_scalars']
    peak, peak_src = measured_peaks()
    dom = int(np.argmax(kms[:3]))
    kern = {}
    for i in range(3):
        gbs = alg_bytes[names[i]] / (kms[i] / 1000.0) / 1e9
        kern[names[i]] = {'ms': kms[i], 'algorithmic_GB': alg_bytes[names[i]] / 1e9,
                          'GBps': gbs, 'frac': gbs / peak}
    kern['k_admm_scalars'] = {'ms': kms[3]}
    traffic, traffic_src = recorded_traffic(names[dom])
    ms_step = ms_max / args.steps
    iter_gb = sum(alg_bytes.values()) / 1e9
    roof = {'bound': 'hbm', 'kernel': names[dom] + (' (fused with the next row-forward)' if sched['fused'] and dom == 2 else ''),
            'achieved': kern[names[dom]]['GBps'],
            'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
            'frac': kern[names[dom]]['frac'], 'traffic': traffic,
            'traffic_source': traffic_src,
            'schedule': sched,
            'iteration_algorithmic_GB': iter_gb,
            # of the timed loop itself (ms_per_step); iterations in which rho changed redo the row-forward
            # pass, which the algorithmic bytes do not credit
            'iteration_frac': iter_gb / (ms_step / 1000.0) / peak,
            'iteration_frac_profiled_phase': iter_gb / (sum(kms) / 1000.0) / peak,
            'kernels': kern}

    parity = None
    if world > 1:
        parity = multi_gpu_parity(dist, rank, world, local_rank)

    # ---- end to end through the public class: host arrays in, coefficient maps out
    # (one untimed pass first: it fills the library's device / pinned allocation pools, which a
    #  long-lived process pays for once)
    wopt = dict(opt)
    wopt['MaxMainIter'] = 3
    bw = cbpdn.ConvBPDN(D, S, LMBDA, cbpdn.ConvBPDN.Options(wopt), dimK=1, device=local_rank)
    if world > 1:
        bw.attach_process_group(dist)
    Yw = bw.solve()
    del bw, Yw
    import gc
    gc.collect()
    barrier()
    t0 = time.perf_counter()
    b2 = cbpdn.ConvBPDN(D, S, LMBDA, cbpdn.ConvBPDN.Options(opt), dimK=1, device=local_rank)
    if world > 1:
        b2.attach_process_group(dist)
    Y = b2.solve()
    rho_final = float(b2.rho)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    te = torch.tensor([t1 - t0], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = args.steps / float(te.item())
    e2e = {'value': e2e_val, 'unit': 'iterations/s',
           'h2d_bytes_per_step': (D.nbytes + S.nbytes) / float(args.steps),
           'd2h_bytes_per_step': (Y.nbytes + 32) / float(args.steps),
           'note': 'construct from host D,S + solve(%d iterations) + getcoef to host; wall clock '
                   'with device synchronisation on both sides' % args.steps}
    del b2

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    if parity is not None and not parity['ok']:
        print(json.dumps({'error': 'multi-GPU parity check failed', 'parity_check': parity}))
        raise SystemExit(3)

    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        cw, cs_ = 3, 5
        per = cpu_solve(K_PER_GPU, cw, cs_, 'scipy', cores)
        cpu = {'value': 1.0 / per, 'unit': 'iterations/s', 'cores': cores, 'kind': 'port',
               'sample': ('oracle (numpy restatement of the reference, bit-identical to it in the build '
                          'container) on the full 32-image batch, scipy.fft workers=%d standing in for the '
                          "reference's multi-threaded pyfftw: %d warm-up + %d timed iterations of one solve, "
                          '%.2f s per iteration' % (cores, cw, cs_, per))}

    cfgs = None
    if world == 1 and not args.no_configs:
        # the other BASELINE.json configurations, short untimed-by-the-headline runs (tools/bench_configs.py)
        sys.path.insert(0, os.path.join(ROOT, 'tools'))
        import bench_configs
        cfgs = {}
        for name in ('cfg2', 'cfg3a', 'cfg3b', 'cfg4', 'cfg5', 'cfg5_cns'):
            try:
                cfgs[name] = bench_configs.measure(name, peak, quick=True)
            except Exception as e:        # a configuration must not take the headline line down
                cfgs[name] = {'error': '%s: %s' % (type(e).__name__, e)}
        try:
            cfgs['modes'] = measure_modes(local_rank, steps=min(args.steps, 50), warm=max(args.warmup, 3))
        except Exception as e:
            cfgs['modes'] = {'error': '%s: %s' % (type(e).__name__, e)}

    out = {
        'metric': METRIC, 'value': value, 'unit': 'iterations/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD,
                   'images_total': world * K_PER_GPU,
                   'image_iterations_per_s': img_iter_per_s,
                   'parallelism': 'images sharded %d per GPU; one allreduce of 5 doubles per iteration'
                                  % K_PER_GPU if world > 1 else 'single GPU',
                   'l2': 'per-iteration working set 1.6 GB per GPU >> 126 MB L2 (no flush needed)',
                   'final_rho': rho_final},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches),
        'roofline': roof,
    }
    if cpu is not None:
        out['cpu_baseline'] = cpu
    if parity is not None:
        out['parity_check'] = parity
    if cfgs is not None:
        out['configs'] = cfgs
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def measure_modes(device, steps=20, warm=5):
    """The other two timing modes of SURVEY.md section 8(d) at the metric configuration (the headline is mode B:
    FastSolve with AutoRho): (A) the class defaults -- residuals, AutoRho and the objective / iteration record every
    iteration (sporco/admm/admm.py:350-377) --, (C) FastSolve without AutoRho -- pure x / y / u steps, no reductions.
    Device time (CUDA events on the library's stream) of `steps` iterations after `warm`."""
    from sporco_b200.admm import cbpdn
    out = {}
    for name, o, rows in (('A_default', {'RelStopTol': 0.0, 'AutoRho': {'Enabled': True}}, True),
                          ('C_fastsolve_fixed_rho', {'RelStopTol': 0.0, 'FastSolve': True,
                                                     'AutoRho': {'Enabled': False}}, False)):
        D, S = make_inputs(K_PER_GPU)
        opt = dict(o)
        opt['MaxMainIter'] = steps
        b = cbpdn.ConvBPDN(D, S, LMBDA, cbpdn.ConvBPDN.Options(opt), dimK=1, device=device)
        h = b._h
        h.admm_configure(**b._admm_config())
        h.admm_iterate(warm, rows)
        _, done, _ = h.admm_iterate(steps, rows)
        ms, _ = h.admm_last_timing()
        out[name] = {'ms_per_step': ms / max(done, 1), 'it_per_s': done / (ms / 1e3), 'steps': int(done), 'warmup': warm}
        del b, h
    return out


def run_cfg5(args):
    """`--config cfg5 | cfg5_cns`: BASELINE.json's configuration 5 -- dictlrn.cbpdndl.ConvBPDNDictLearn on 16 images
    256x256 with an 8x8x64 dictionary, ADMM X step and PGM (`cfg5`) or consensus ADMM (`cfg5_cns`: "alternating X/D
    ADMM" as the configuration is written) D step -- with the 16 training images sharded over the ranks (strong
    scaling: 16 / N images per GPU).  A "step" is one outer iteration; wall clock around `solve()` with device
    synchronisation and a barrier on both sides, max over ranks.  At N > 1 an untimed small sharded run is compared
    with the same run on one GPU (all images) and the run fails if they disagree."""
    import torch
    from sporco_b200.dictlrn import cbpdndl
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    dist = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    dmethod = 'cns' if args.config == 'cfg5_cns' else 'pgm'
    ccmod = {'rho': 16.0} if dmethod == 'cns' else {}
    rng = np.random.default_rng(2024)
    D0 = rng.standard_normal((8, 8, 64)).astype(np.float32)
    S = rng.standard_normal((256, 256, 16)).astype(np.float32)
    if 16 % world:
        raise SystemExit('the 16 training images must divide over the ranks')
    per = 16 // world
    mine = list(range(rank * per, (rank + 1) * per))

    def learner(D0_, S_, iters, dev):
        o = cbpdndl.ConvBPDNDictLearn.Options({'MaxMainIter': iters, 'CCMOD': dict(ccmod)}, dmethod=dmethod)
        return cbpdndl.ConvBPDNDictLearn(D0_, S_, 0.1, o, dmethod=dmethod, device=dev)

    parity = None
    if world > 1:
        Ds = rng.standard_normal((6, 6, 8)).astype(np.float32)
        Ss = rng.standard_normal((64, 64, 2 * world)).astype(np.float32)
        ps = learner(Ds, np.ascontiguousarray(Ss[:, :, 2 * rank:2 * rank + 2]), 12, local_rank)
        ps.attach_process_group(dist)
        Dsh = ps.solve().squeeze()
        obj_sh = np.array(ps.getitstat().ObjFun, dtype=np.float64)
        res = torch.zeros(2, dtype=torch.float64, device='cuda')
        if rank == 0:
            p1 = learner(Ds, Ss, 12, local_rank)
            D1 = p1.solve().squeeze()
            obj1 = np.array(p1.getitstat().ObjFun, dtype=np.float64)
            res[0] = float(np.linalg.norm((Dsh - D1).ravel()) / np.linalg.norm(D1.ravel()))
            res[1] = float(np.max(np.abs(obj_sh - obj1) / np.abs(obj1)))
            del p1
        dist.broadcast(res, src=0)
        rel_d, rel_obj = float(res[0].item()), float(res[1].item())
        parity = {'n': world, 'problem': '64x64, 6x6x8 dictionary, %d images (2 per rank), 12 outer iterations, sharded '
                  'against one GPU holding all images' % (2 * world), 'rel_D': rel_d, 'rel_ObjFun': rel_obj,
                  'tol': 3e-4, 'ok': bool(rel_d < 3e-4 and rel_obj < 1e-4)}
        del ps
    b = learner(D0, np.ascontiguousarray(S[:, :, mine]), max(args.warmup, 3), local_rank)
    if world > 1:
        b.attach_process_group(dist)
    b.solve()
    b.opt['MaxMainIter'] = args.steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t0 = time.perf_counter()
    b.solve()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    clocks = sampler.stop() if rank == 0 else None
    te = torch.tensor([t1 - t0], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    sec = float(te.item())
    its = b.getitstat()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    if parity is not None and not parity['ok']:
        print(json.dumps({'error': 'multi-GPU parity check failed', 'parity_check': parity}))
        raise SystemExit(3)
    out = {'metric': 'ConvBPDNDictLearn outer iterations/sec (16 images 256x256, 8x8x64 dictionary, float32)',
           'value': args.steps / sec, 'unit': 'iterations/s', 'n_gpus': world, 'steps': args.steps,
           'warmup': max(args.warmup, 3), 'ms_per_step': 1e3 * sec / args.steps, 'higher_is_better': True,
           'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
           'config': {'workload': 'dictlrn.cbpdndl.ConvBPDNDictLearn, ADMM X step / %s D step, 16 training images '
                                  'sharded %d per GPU; per outer iteration the ranks exchange the 8x8x64 filter supports '
                                  '(16 KB) and the residual / objective sums over peer memory'
                                  % ('consensus ADMM' if dmethod == 'cns' else 'PGM', per),
                      'timing': 'wall clock around solve() with device synchronisation, max over ranks (the host loop '
                                'reads one record of scalars per outer iteration: part of the algorithm)',
                      'ObjFun_first_last': [float(its.ObjFun[0]), float(its.ObjFun[-1])]},
           'clocks': clocks}
    if parity is not None:
        out['parity_check'] = parity
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-configs', action='store_true', help='skip the configs block (cfg2..cfg5)')
    ap.add_argument('--config', default='metric', choices=['metric', 'cfg5', 'cfg5_cns'],
                    help="'metric' (default): the headline ConvBPDN benchmark; cfg5 / cfg5_cns: dictionary learning with "
                         "the training images sharded over the GPUs")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    elif args.config != 'metric':
        run_cfg5(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
