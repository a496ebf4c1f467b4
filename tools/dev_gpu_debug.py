import sys, time, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle')
from ceit_b200 import _lib
import ceit_oracle as o
torch.manual_seed(0)
dev = 'cuda'
print(torch.cuda.get_device_name(0))
# --- dgemm
for (M, N, K) in [(64, 64, 16), (100, 37, 53), (273, 264, 202), (1000, 520, 333)]:
    for ta in (False, True):
        for tb in (False, True):
            A = torch.randn((K, M) if ta else (M, K), dtype=torch.float64, device=dev)
            B = torch.randn((N, K) if tb else (K, N), dtype=torch.float64, device=dev)
            C0 = torch.randn((M, N), dtype=torch.float64, device=dev)
            C = C0.clone()
            _lib.dgemm(A, B, ta, tb, alpha=0.7, beta=-0.3, C=C)
            ref = 0.7 * ((A.T if ta else A) @ (B.T if tb else B)) - 0.3 * C0
            err = (C - ref).abs().max().item() / ref.abs().max().item()
            print('dgemm', M, N, K, ta, tb, 'relerr', err)
for tag in ('21_21', '50_50'):
    m = dict(np.load(f'tests/golden/mesh_{tag}.npz')); r = dict(np.load(f'tests/golden/ref_{tag}.npz'))
    ptr, ee = m['electrode_ptr'], m['electrode_elems']
    els = [ee[ptr[i]:ptr[i + 1]] for i in range(16)]
    nodes, elem = m['nodes'], m['elem']; N, E, L = len(nodes), len(elem), 16
    omega = 2 * np.pi * float(r['signal_frequency'])
    for tb_ in (0, 64, 100000):
        p = _lib.Plan(nodes, elem, els, target_block=tb_)
        print(tag, 'plan', p.info()[:12])
        perm = torch.tensor(m['elem_perm'], device=dev); var = torch.zeros(E, dtype=torch.float64, device=dev)
        vals = torch.zeros((p.nnz, 2), dtype=torch.float64, device=dev)
        ep = torch.zeros((E, 9), dtype=torch.float64, device=dev)
        p.assemble(perm, var, omega, vals, ep)
        torch.cuda.synchronize()
        v = vals.cpu().numpy(); vc = v[:, 0] + 1j * v[:, 1]
        print(' K relerr', np.abs(vc - r['K_data']).max() / np.abs(r['K_data']).max(), 'elem_param err', np.abs(ep.cpu().numpy() - m['elem_param']).max())
        t0 = time.time(); p.factor(False); torch.cuda.synchronize(); print(' factor s', time.time() - t0, 'flag', p.numeric_flag())
        J = torch.zeros((L * (L - 1), E), dtype=torch.float64, device=dev); base = torch.zeros(L * (L - 1), dtype=torch.float64, device=dev)
        t0 = time.time(); p.jacobian_block(0, L, 0, E, 1e-3, omega, J, E, base); torch.cuda.synchronize(); print(' jac s', time.time() - t0, 'flag', p.numeric_flag())
        Jh = J.cpu().numpy()
        print(' base-1', np.abs(base.cpu().numpy() - 1).max(), 'J range', (Jh - 1).min(), (Jh - 1).max())
        sp = r['sample_pairs']
        d = max(np.abs(Jh[i * 15:(i + 1) * 15, j] - r['sample_cols'][k]).max() for k, (i, j) in enumerate(sp))
        print(' sample cols maxabs', d)
        if 'JAC' in r:
            print(' full J maxabs', np.abs(Jh - r['JAC']).max(), 'J-1 fro rel', np.linalg.norm(Jh - r['JAC']) / np.linalg.norm(r['JAC'] - 1))
        # timing second run
        torch.cuda.synchronize(); t0 = time.time()
        p.assemble(perm, var, omega); p.factor(False); p.jacobian_block(0, L, 0, E, 1e-3, omega, J, E, base); torch.cuda.synchronize()
        print(' full pipeline 2nd run s', time.time() - t0, 'launches', _lib.launch_count())
        # forward
        na = torch.zeros(N, dtype=torch.float64, device=dev); epot = torch.zeros(E, dtype=torch.float64, device=dev); el = torch.zeros(L, dtype=torch.float64, device=dev)
        p.forward(3, na, epot, el); torch.cuda.synchronize()
        print(' forward node diff', np.abs(na.cpu().numpy() - r['base_node_potential'][3]).max(), np.abs(el.cpu().numpy() - r['base_electrode_potential'][3].real).max())
        # complex path: object
        var_o = torch.tensor(r['obj_variable'], device=dev)
        p.assemble(perm, var_o, omega); p.factor(True); p.forward(2, na, epot, el); torch.cuda.synchronize()
        print(' complex forward node diff', np.abs(na.cpu().numpy() - r['obj_node_potential']).max(), np.abs(el.cpu().numpy() - r['obj_electrode_potential'].real).max(), 'flag', p.numeric_flag())
        if 'nz_sample_pairs' in r:
            var_n = torch.full((E,), float(r['nz_base_variable']), dtype=torch.float64, device=dev)
            p.assemble(perm, var_n, omega); p.factor(True); p.jacobian_block(0, L, 0, E, 1e-3, omega, J, E, base); torch.cuda.synchronize()
            Jh = J.cpu().numpy()
            d = max(np.abs(Jh[i * 15:(i + 1) * 15, j] - r['nz_sample_cols'][k]).max() for k, (i, j) in enumerate(r['nz_sample_pairs']))
            print(' complex base sample cols maxabs', d)
        p.close()
    if 'JAC' in r:
        Jd = torch.tensor(r['JAC'], device=dev); det = torch.tensor(m['detection_index'], device=dev)
        for form in (1, 2):
            for lam, key in ((289, 'inv_mat_289'), (2e-5, 'inv_mat_2e-5')):
                inv = _lib.inverse_model(Jd, det, lam, form=form); torch.cuda.synchronize()
                print(' inverse model form', form, 'lambda', lam, 'rel', np.abs(inv.cpu().numpy() - r[key]).max() / np.abs(r[key]).max())
        inv = _lib.inverse_model(Jd, det, 289)
        dV = torch.tensor(r['solve_dV'], device=dev)
        out = _lib.reconstruct(inv, dV); torch.cuda.synchronize()
        print(' solve rel', np.abs(out.cpu().numpy() - r['solve_out']).max() / np.abs(r['solve_out']).max())
        dVb = torch.rand((240, 1000), dtype=torch.float64, device=dev)
        outb = _lib.reconstruct(inv, dVb); torch.cuda.synchronize()
        print(' batched solve rel', ((outb - inv @ dVb).abs().max() / outb.abs().max()).item())
