import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from financial_rag_system_b200.checkpoint import MINILM_L6_CE, BGE_SMALL, synthetic_checkpoint
from financial_rag_system_b200.encoder import BertEncoder
from oracle import encoder_oracle as eo
sys.path.insert(0, "tests")
from test_encoder_gpu import _random_batch
w = synthetic_checkpoint(MINILM_L6_CE, 4321)
enc = BertEncoder(MINILM_L6_CE, w, device=0, max_tokens=8192)
rng = np.random.default_rng(2)
lens = [int(x) for x in rng.integers(20, 513, size=15)] + [512, 4]
ids, tts, cu = _random_batch(lens, 21, pairs=True)
got = enc.score_packed(ids, tts, cu)
ref = eo.score_pairs(MINILM_L6_CE, w, ids, tts, cu)
d = got - ref
print("lib", os.environ.get("FRS_B200_LIB", "default"))
print("signed err: mean %.4f std %.4f  max|.| %.4f" % (d.mean(), d.std(), np.abs(d).max()))
print(np.round(d, 3))
hid = enc.last_hidden(int(cu[-1])).cpu().numpy()
ref_h = eo.last_hidden_packed(MINILM_L6_CE, w, ids, tts and cu, tts) if False else None
enc.close()
