"""diagnostic (not a test): error breakdown of the tcgen05 train step by expert tile / team tile"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np, torch
from opentf_b200 import ops
from test_gpu_tc import run_tc, make_case
from test_gpu_kernels import oracle_out, run_out_train
ws = ops.Workspace(torch.device('cuda:0'))
for (B, E) in [(64, 128), (64, 256), (128, 128), (100, 200), (777, 5000)]:
    A, W, b, Y, negs = make_case(B, E, 7 * B + E)
    loss, dW, db, dA, _ = run_tc(ops, ws, A, W, b, Y, negs, 10.0, 1.0)
    l_ref, dW_ref, db_ref, dA_ref = oracle_out(A, W, b, Y, negs, 10.0, 1.0)
    print(f'--- B={B} E={E} loss {loss:.5f} ref {l_ref:.5f}')
    e_db = (db - db_ref).abs().numpy(); print(' db  max err', e_db.max(), 'ref max', db_ref.abs().max().item(), 'worst experts', np.argsort(-e_db)[:8], 'their err', np.sort(e_db)[::-1][:8])
    bad = np.nonzero(e_db > 1e-3 * db_ref.abs().max().item())[0]
    print(' db  #bad', len(bad), 'bad%128', np.unique(bad % 128)[:20], 'bad//128', np.unique(bad // 128)[:20])
    if len(bad):
        j = bad[0]; print('   expert', j, 'db', db[j].item(), 'ref', db_ref[j].item(), 'members col sum', int(np.asarray(Y[:, j].sum())), 'neg hits', int((negs == j).sum()), 'rows with member', Y[:, j].nonzero()[0][:10], 'rows with neg', np.nonzero((negs == j).any(1))[0][:10])
    e_dw = (dW - dW_ref).abs().numpy(); print(' dW  max err', e_dw.max(), 'ref max', dW_ref.abs().max().item(), 'rows with big err', np.unique(np.nonzero(e_dw > 4e-3 * dW_ref.abs().max().item())[0])[:10])
    e_da = (dA - dA_ref).abs().numpy(); print(' dA  max err', e_da.max(), 'ref max', dA_ref.abs().max().item(), 'rows with big err', np.unique(np.nonzero(e_da > 4e-3 * dA_ref.abs().max().item())[0])[:10])
