"""Two host threads driving the same models on their own CUDA streams at the same time (the reference's SLAM runs the model from a
worker thread while the GUI thread owns the main one, slam.py:533): results must equal the single-threaded ones."""
import os, sys, threading
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from must3r_b200 import synthetic as syn  # noqa: E402
from must3r_b200.model import Dust3rEncoder, MUSt3R, set_precision  # noqa: E402

set_precision(torch.float16)
enc = Dust3rEncoder(img_size=(224, 224), depth=6); dec = MUSt3R(img_size=(224, 224), depth=6, feedback_type="single_mlp", memory_mode="kv", landscape_only=False)
enc.load_state_dict(syn.encoder_state_dict(3, depth=6)); dec.load_state_dict(syn.decoder_state_dict(3, depth=6))
enc, dec = enc.cuda().eval(), dec.cuda().eval()


def chain(seed, out, use_stream):
    st = torch.cuda.Stream() if use_stream else torch.cuda.current_stream()
    with torch.cuda.stream(st):
        imgs, ts = syn.synthetic_views(6, 224, 224, seed=seed)
        imgs, ts = imgs.cuda(), ts.cuda()
        x, pos = enc(imgs, ts)
        mem, _ = dec(x[None, :2], pos[None, :2], ts[None, :2], None)
        for i in range(2, 6):
            mem, pm = dec(x[None, i:i + 1], pos[None, i:i + 1], ts[None, i:i + 1], mem)
        _, pm = dec(x[None], pos[None], ts[None], mem, render=True)
        st.synchronize()
        out[seed] = pm.clone()


ref = {}
for s in (1, 2):
    chain(s, ref, False)
ok = True
for rep in range(5):
    got = {}
    th = [threading.Thread(target=chain, args=(s, got, True)) for s in (1, 2)]
    [t.start() for t in th]; [t.join() for t in th]
    ok = ok and all(torch.equal(got[s], ref[s]) for s in (1, 2))
print("two threads on two streams == single thread:", ok)
sys.exit(0 if ok else 1)
