"""Scratch GPU check used during bring-up: parity vs oracle on a few configs + headline timing."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import torch
from oracle import oracle as O
from ldpc_3gpp_matlab_b200 import capi

def make_llr(bg, Z, B, E, esn0, rng, filler=0):
    d = O.dims(bg, Z)
    info = rng.integers(0, 2, (B, d['K']), dtype=np.uint8)
    if filler: info[:, d['K']-filler:] = 0
    cw = O.encode(bg, Z, info)
    s2 = 10 ** (-esn0 / 10)
    y = (1 - 2.0 * cw) / np.sqrt(2) + rng.normal(0, np.sqrt(s2 / 2), cw.shape)
    llr = (2 * np.sqrt(2) * y / s2).astype(np.float32)
    llr[:, :2 * Z] = 0
    llr[:, 2 * Z + E:] = 0
    if filler: llr[:, d['K']-filler:d['K']] = np.inf
    return info, llr

rng = np.random.default_rng(1)
res = {}
for (bg, Z, B, E, esn0, iters, et, rows, fill) in [
    (1, 384, 8, 25272, -0.5, 8, False, 0, 0),
    (1, 384, 8, 9478, 6.0, 20, True, 5, 0),
    (2, 52, 40, 2000, -2.0, 8, False, 33, 104),
    (2, 6, 300, 100, 0.0, 8, True, 13, 24),
    (1, 2, 500, 100, 2.0, 8, True, 0, 0),
    (2, 13, 100, 400, 0.0, 10, True, 0, 0),
    (1, 208, 6, 208*40, 1.0, 8, False, 0, 0),
]:
    info, llr = make_llr(bg, Z, B, E, esn0, rng, fill)
    ref = O.decode_nms(bg, Z, llr, iters, early_term=et, n_rows=rows)
    h = capi.Handle(bg, Z, iters, et)
    out = h.decode(llr, n_rows=rows, want_soft=True)
    same_hard = bool((out['hard'] == ref['hard']).all())
    same_app = bool((out['app'].view(np.uint32) == ref['app'].view(np.uint32)).all())
    same_it = bool((out['iters'] == ref['iters']).all())
    same_ok = bool((out['parity_ok'] == ref['parity_ok']).all())
    print(bg, Z, B, 'hard', same_hard, 'app', same_app, 'iters', same_it, 'ok', same_ok,
          'bler', float((out['hard'] != info).any(1).mean()), 'mean_it', float(out['iters'].mean()), flush=True)
    res[f'{bg}_{Z}_{E}'] = [same_hard, same_app, same_it, same_ok]
    h.close()

# headline timing, device-resident
bg, Z, B = 1, 384, 4096
d = O.dims(bg, Z)
info, llr1 = make_llr(bg, Z, 64, 25272, -0.3, rng)
llr = torch.from_numpy(np.tile(llr1, (B // 64, 1))).cuda()
hard = torch.zeros((B, d['K']), dtype=torch.uint8, device='cuda')
h = capi.Handle(bg, Z, 8, False)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    h.decode_raw(llr, B, hard, mem=capi.MEM_DEVICE, stream=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5):
    h.decode_raw(llr, B, hard, mem=capi.MEM_DEVICE, stream=st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('headline ms', ms, 'Gb/s', B * d['K'] / ms / 1e6)
hh = hard.cpu().numpy()
print('bler headline', float((hh[:64] != info).any(1).mean()))
res['headline_ms'] = ms
json.dump(res, open('gpurun_out/first.json', 'w'))
