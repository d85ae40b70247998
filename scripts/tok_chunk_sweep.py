"""Text head time against the token chunk size (T2L_TOK_CHUNK): wave quantisation of the four token GEMMs and L2 residency of
the activations between them.  One engine per setting, interleaved repeats."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth
from text2loc_b200.engine import Engine

sd = synth.make_state_dict(0)
t5 = torch.from_numpy(synth.make_t5_features(2, 4096)).cuda()
t16 = t5.half()
chunks = [int(a) for a in sys.argv[1:]] or [9472, 18944, 28416, 32768, 37888, 75776]
engines = {}
for c in chunks:
    os.environ["T2L_TOK_CHUNK"] = str(c)
    engines[c] = Engine("cuda:0")
    engines[c].load_state_dict(sd)
    for _ in range(2):
        engines[c].encode_text(t5, 6); engines[c].encode_text(t16, 6)
res = {(c, n): [] for c in chunks for n in ("f32", "f16")}
for rep in range(4):
    for c in chunks:
        for name, x in (("f32", t5), ("f16", t16)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                engines[c].encode_text(x, 6)
            b.record(); torch.cuda.synchronize()
            res[(c, name)].append(a.elapsed_time(b) / 3)
for c in chunks:
    print(c, {n: [round(v, 3) for v in res[(c, n)]] for n in ("f32", "f16")})
