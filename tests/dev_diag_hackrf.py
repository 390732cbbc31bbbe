import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
from scanner_b200 import SpectrumSense
G = np.load("tests/golden/hackrf_vectors.npz")
N, FS, START, STOP, THR, IT, VALID, TPS = [G["sweep_params"][i] for i in range(8)]
N, FS, VALID = int(N), int(FS), int(VALID)
stream = G["sweep_stream"]; nt = stream.shape[0]; chunks = VALID // (2 * N)
window, use_w = O.window_build(5, N), O.use_window(0.75, N)
patched, ofreq, _ = O.hackrf_prepass(stream, VALID)
res = O.pipeline(patched.view(np.int8).reshape(-1, N, 2), N, FS, 8, 1, True, 1, float(THR), window, use_w, precision=1, want_f64=True)
ss = SpectrumSense(sample_count=N, sample_rate=FS, enob=8, sample_kind=1, correct_dc_offset=True, threshold=float(THR), window=window, max_spectra=nt * chunks)
host = ss.process(patched.view(np.int8).reshape(-1, N, 2))
print("host path mask equal:", np.array_equal(host["hit_mask"], res["hit_mask"]))
d = torch.from_numpy(stream.copy()).cuda()
masks = torch.zeros(nt * chunks, ss.words, dtype=torch.int32, device="cuda")
counts = torch.zeros(nt * chunks, dtype=torch.int32, device="cuda")
spec = torch.zeros(nt * chunks, N, dtype=torch.float32, device="cuda")
ss.hackrf_prepass_device(d.data_ptr(), nt, VALID, 0, 0, 0)
ss.launch_device(d.data_ptr(), nt * chunks, spec.data_ptr(), masks.data_ptr(), counts.data_ptr(), 0, 0, 0)
torch.cuda.synchronize()
m = masks.cpu().numpy().view(np.uint32)
print("device bytes equal:", np.array_equal(d.cpu().numpy(), patched))
bad = np.nonzero((m != res["hit_mask"]).any(axis=1))[0]
print("bad rows", bad.tolist())
sp = spec.cpu().numpy()
for b in bad[:6]:
    x = m[b] ^ res["hit_mask"][b]
    for w in np.nonzero(x)[0]:
        for bit in range(32):
            if x[w] >> bit & 1:
                i = w * 32 + bit; j = (i + N // 2) % N
                print(b, "i", i, "j", j, "gpu dB", sp[b, j], "oracle f32", res["spectra_db"][b, j], "f64", res["spectra_db64"][b, j], "thr", THR,
                      "rms dB", 10*np.log10(np.mean(10**(res["spectra_db64"][b]/5)))/2)
