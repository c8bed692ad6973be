"""One memory-update step (1 view 512x384 against M=10 memory views) and one render call (8 views) inside a
cudaProfilerStart/Stop range, for `ncu --profile-from-start off` launch lists.  Also prints event-timed step latencies."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from must3r_b200 import synthetic as syn
from must3r_b200.model import Dust3rEncoder, MUSt3R, set_precision

set_precision(torch.bfloat16)
enc = Dust3rEncoder(img_size=(512, 512)); dec = MUSt3R(img_size=(512, 512), feedback_type="single_mlp", memory_mode="kv", landscape_only=False)
enc.load_state_dict(syn.encoder_state_dict(0)); dec.load_state_dict(syn.decoder_state_dict(0))
enc, dec = enc.cuda().eval(), dec.cuda().eval()
V = 12
imgs, ts = syn.synthetic_views(V, 384, 512, seed=2); imgs, ts = imgs.cuda(), ts.cuda()
from must3r_b200.engine.inference import _with_host_shape
_ts_host = ts.cpu()
x, pos = enc(imgs, ts)
mem, _ = dec(x[None, :2], pos[None, :2], ts[None, :2], None)
for i in range(2, 10):
    mem, _ = dec(x[None, i:i + 1], pos[None, i:i + 1], ts[None, i:i + 1], mem)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
# timed (unprofiled) repeats of the same step for a latency number
for rep in range(3):
    ev[0].record(); m2, _ = dec(x[None, 10:11], pos[None, 10:11], _with_host_shape(_ts_host[10:11], 'cuda'), mem); ev[1].record()
torch.cuda.synchronize()
print(f"update step (1 view, M=10): {ev[0].elapsed_time(ev[1]):.3f} ms", flush=True)
import time
torch.cuda.synchronize()
_tsd = _with_host_shape(_ts_host[10:11], 'cuda')
t0 = time.perf_counter()
for rep in range(5):
    m2, _ = dec(x[None, 10:11], pos[None, 10:11], _with_host_shape(_ts_host[10:11], 'cuda'), mem)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
# split the host time: C call (kernel enqueue from C++) vs Python wrapper around it
from must3r_b200 import _lib as _L
_orig = _L.lib().m3r_decoder_forward
_acc = [0.0, 0]
def _timed(*a):
    t = time.perf_counter(); r = _orig(*a); _acc[0] += time.perf_counter() - t; _acc[1] += 1; return r
class _Proxy:
    def __getattr__(self, n): return _timed if n == "m3r_decoder_forward" else getattr(_L._lib, n)
_real = _L.lib
_L.lib = lambda: _Proxy()
import must3r_b200.model.decoder as _D
_D._lib.lib = _L.lib
for rep in range(5):
    m2, _ = dec(x[None, 10:11], pos[None, 10:11], _with_host_shape(_ts_host[10:11], 'cuda'), mem)
torch.cuda.synchronize()
_L.lib = _real; _D._lib.lib = _real
print(f"  of which inside m3r_decoder_forward (C++ enqueue of ~220 launches): {_acc[0] / max(_acc[1], 1) * 1e3:.3f} ms", flush=True)
print(f"host time per update call (enqueue only): {(t1 - t0) / 5 * 1e3:.3f} ms; with final sync: {(t2 - t0) / 5 * 1e3:.3f} ms", flush=True)
for rep in range(2):
    ev[2].record(); dec(x[None, :8], pos[None, :8], _with_host_shape(_ts_host[:8], 'cuda'), mem, render=True); ev[3].record()
torch.cuda.synchronize()
print(f"render (8 views, M=10): {ev[2].elapsed_time(ev[3]):.3f} ms", flush=True)
torch.cuda.cudart().cudaProfilerStart()
m2, _ = dec(x[None, 10:11], pos[None, 10:11], _with_host_shape(_ts_host[10:11], 'cuda'), mem)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
