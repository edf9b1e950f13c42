"""Device-resident batches (kp_tokenize_batch_device8), both paths, CUDA-event time per call at several sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import kanpyo_b200
from kanpyo_b200 import builder, corpus
d = builder.ipadic()
vocab = corpus.Vocabulary(d.keywords, d.morphs)
text, off = corpus.synth_corpus(vocab, 65536, "cfg2")
tk = kanpyo_b200.Tokenizer(d, device=0)
for n in (256, 512, 1024, 2048, 3072, 4096, 5120, 6144, 8192):
    t_ = torch.from_numpy(text[:int(off[n])].copy()).cuda()
    o_ = torch.from_numpy(off[:n + 1].astype(np.int64)).cuda()
    torch.cuda.synchronize()
    row = {}
    for path in ("pipeline", "fused"):
        tk.set_path(path)
        for _ in range(5):
            tk.tokenize_batch_device8(t_.data_ptr(), o_.data_ptr(), n, 0, int(off[n]))
        ms = []
        for _ in range(20):
            tk.tokenize_batch_device8(t_.data_ptr(), o_.data_ptr(), n, 0, int(off[n]))
            ms.append(tk.profile()["total_ms"])
        ms.sort()
        row[path] = ms[len(ms) // 2]
    print(n, {k: round(v, 3) for k, v in row.items()}, flush=True)
