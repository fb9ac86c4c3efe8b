"""Run under torchrun with 2 (or more) ranks (one GPU each): sporco_b200 with the batch sharded over
the ranks against the single-object oracle on the whole batch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import cbpdn_oracle as orc
    from sporco_b200.admm import cbpdn
    rank = int(os.environ['RANK'])
    world = int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rng = np.random.default_rng(77)
    D = rng.standard_normal((8, 8, 16)).astype(np.float32)
    S = rng.standard_normal((128, 128, 3 * world if world <= 2 else 2 * world)).astype(np.float32)
    per = S.shape[2] // world
    mine = list(range(rank * per, (rank + 1) * per))
    opt = {'MaxMainIter': 20, 'RelStopTol': 0.0}
    b = cbpdn.ConvBPDN(D, S[:, :, mine], 0.1, cbpdn.ConvBPDN.Options(opt), dimK=1, device=local)
    b.attach_process_group(dist)
    Y = b.solve()
    r = orc.admm_convbpdn(D, S, 0.1, opt=opt, dimK=1)
    err = np.linalg.norm((Y - r.Y[:, :, :, mine, :]).ravel()) / np.linalg.norm(r.Y[:, :, :, mine, :].ravel())
    rho = np.array(b.getitstat().Rho, dtype=np.float64)
    rho_ref = np.array([row[8] for row in r.itstat], dtype=np.float64)
    rho_err = np.abs(rho - rho_ref).max() / np.abs(rho_ref).max()
    obj = np.array(b.getitstat().ObjFun, dtype=np.float64)
    obj_ref = np.array([row[1] for row in r.itstat], dtype=np.float64)
    obj_err = np.abs(obj - obj_ref).max() / np.abs(obj_ref).max()
    ok = err < 3e-4 and rho_err < 1e-4 and obj_err < 1e-4
    # dictionary learning with the training images sharded: gradient summed over ranks on the device
    from oracle import cbpdndl_oracle as ocdl
    from sporco_b200.dictlrn import cbpdndl
    D0 = rng.standard_normal((6, 6, 8)).astype(np.float32)
    o = {'MaxMainIter': 15, 'CBPDN': {'rho': 5.0, 'AutoRho': {'Period': 4}}, 'CCMOD': {'L': 60.0}}
    dl = cbpdndl.ConvBPDNDictLearn(D0, S[:, :, mine], 0.1, cbpdndl.ConvBPDNDictLearn.Options(o),
                                   device=local)
    dl.attach_process_group(dist)
    Dl = dl.solve().squeeze()
    rd = ocdl.cbpdndl(D0, S, 0.1, o)
    d_err = np.linalg.norm((Dl - rd['D']).ravel()) / np.linalg.norm(rd['D'].ravel())
    its = dl.getitstat()
    dobj_err = np.abs(np.array(its.ObjFun) - rd['ObjFun']).max() / np.abs(rd['ObjFun']).max()
    drs_err = np.abs(np.array(its.D_Rsdl) - rd['D_Rsdl']).max() / np.abs(rd['D_Rsdl']).max()
    print('rank %d: dictlearn D err %.3e obj err %.3e rsdl err %.3e' % (rank, d_err, dobj_err, drs_err),
          flush=True)
    ok = ok and d_err < 1e-4 and dobj_err < 1e-4 and drs_err < 1e-3
    # PGM / FISTA with backtracking, images sharded: the sums behind F <= Q, the residual and the objective are
    # reduced over the ranks, so every rank takes the same backtracking decisions as the single-object solver
    from sporco_b200.pgm import cbpdn as pcbpdn
    from sporco_b200.pgm.backtrack import BacktrackStandard
    po = {'MaxMainIter': 15, 'RelStopTol': 0.0, 'L': 5.0}
    pb = pcbpdn.ConvBPDN(D, S[:, :, mine], 0.1,
                         pcbpdn.ConvBPDN.Options(dict(po, Backtrack=BacktrackStandard(gamma_u=1.3, maxiter=10))),
                         dimK=1, device=local)
    pb.attach_process_group(dist)
    Xp = pb.solve()
    rp = orc.pgm_convbpdn(D, S, 0.1, opt=dict(po, Backtrack={'gamma_u': 1.3, 'maxiter': 10}), dimK=1)
    pits = pb.getitstat()
    x_err = np.linalg.norm((Xp - rp.X[:, :, :, mine, :]).ravel()) / np.linalg.norm(rp.X[:, :, :, mine, :].ravel())
    bt_same = np.array_equal(np.asarray(pits.IterBTrack, dtype=float), np.array([row[7] for row in rp.itstat], dtype=float))
    l_err = np.abs(np.array(pits.L, dtype=np.float64) - np.array([row[8] for row in rp.itstat])).max()
    pobj_err = np.abs(np.array(pits.ObjFun) - np.array([row[1] for row in rp.itstat])).max() / \
        np.abs(np.array([row[1] for row in rp.itstat])).max()
    print('rank %d: pgm X err %.3e backtracking counts equal %s L err %.3e obj err %.3e'
          % (rank, x_err, bt_same, l_err, pobj_err), flush=True)
    ok = ok and x_err < 1e-4 and bt_same and l_err < 1e-4 and pobj_err < 1e-4
    # consensus dictionary update, blocks (images) sharded: supports of the block mean + norms summed over ranks
    from sporco_b200.admm import ccmod
    Zc = rng.standard_normal((64, 64, 1, 2 * world, 6)).astype(np.float32)
    Zc[np.abs(Zc) < 1.0] = 0
    Sc = rng.standard_normal((64, 64, 2 * world)).astype(np.float32)
    cm = [2 * rank, 2 * rank + 1]
    co_ = {'MaxMainIter': 12, 'rho': 2.0, 'AutoRho': {'Enabled': True, 'Period': 3, 'AutoScaling': True, 'Scaling': 10.0}}
    cc = ccmod.ConvCnstrMOD_Consensus(Zc[:, :, :, cm, :], Sc[:, :, cm], (5, 5, 6), ccmod.ConvCnstrMOD_Consensus.Options(co_),
                                      device=local)
    cc.attach_process_group(dist)
    Yc = cc.solve()
    rc = ocdl.ConsensusCCMOD(Sc.astype(np.float64), (5, 5, 6), co_)
    rc.setcoef(Zc.astype(np.float64))
    rc.solve()
    cref = np.array(rc.itstat, dtype=np.float64)
    cits = cc.getitstat()
    cy_err = np.linalg.norm((Yc - rc.Y).ravel()) / np.linalg.norm(rc.Y.ravel())
    crho_err = np.abs(np.array(cits.Rho, dtype=np.float64) - cref[:, 7]).max() / np.abs(cref[:, 7]).max()
    cr_err = np.abs(np.array(cits.PrimalRsdl, dtype=np.float64) - cref[:, 3]).max() / np.abs(cref[:, 3]).max()
    print('rank %d: consensus ccmod Y err (vs float64 oracle) %.3e rho err %.3e r err %.3e' % (rank, cy_err, crho_err, cr_err),
          flush=True)
    ok = ok and cy_err < 1e-4 and crho_err < 1e-4 and cr_err < 1e-3
    # ... and as the D step of dictionary learning with the training images sharded
    o2 = {'MaxMainIter': 12, 'CBPDN': {'rho': 5.0}, 'CCMOD': {'rho': 2.0, 'ZeroMean': True}}
    d1 = cbpdndl.ConvBPDNDictLearn(D0, S[:, :, mine], 0.1, cbpdndl.ConvBPDNDictLearn.Options(o2, dmethod='cns'),
                                   dmethod='cns', device=local)
    d1.attach_process_group(dist)
    D1 = d1.solve().squeeze()
    if rank == 0:
        d2 = cbpdndl.ConvBPDNDictLearn(D0, S, 0.1, cbpdndl.ConvBPDNDictLearn.Options(o2, dmethod='cns'),
                                       dmethod='cns', device=local)
        D2 = d2.solve().squeeze()
        dd_err = np.linalg.norm((D1 - D2).ravel()) / np.linalg.norm(D2.ravel())
        do_err = np.abs(np.array(d1.getitstat().ObjFun) - np.array(d2.getitstat().ObjFun)).max() / \
            np.abs(np.array(d2.getitstat().ObjFun)).max()
        print('rank 0: dictlearn (consensus D step) sharded vs one GPU: D err %.3e obj err %.3e' % (dd_err, do_err), flush=True)
        ok = ok and dd_err < 3e-4 and do_err < 1e-4
    print('rank %d: schedule %s, peer-memory exchange %s' % (rank, b._h.admm_schedule_info(), b._p2p), flush=True)
    t = torch.tensor([1.0 if ok else 0.0], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    print('rank %d: Y err %.3e rho err %.3e obj err %.3e' % (rank, err, rho_err, obj_err), flush=True)
    if rank == 0 and t.item() == 1.0:
        print('MULTI_GPU_OK', flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == '__main__':
    main()
