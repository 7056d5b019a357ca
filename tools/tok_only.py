#!/usr/bin/env python
"""Kernel-variant timing: run only k_tokenize on the C2 text (G2N_DBG_TOKENIZE_ONLY), print its time.
    G2N_LIB=build/libs/libg2n_X.so python tools/tok_only.py"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from bench import make_text  # noqa: E402
from gfa2network_b200 import _capi  # noqa: E402

cfg, text, _, _ = make_text(sys.argv[1] if len(sys.argv) > 1 else "C2")
t = torch.from_numpy(text).cuda()
h = _capi.Handle(0)
p = _capi.Params(0, 0, 0, 0, 0, 0, _capi.FMT_CSR, 1, None, 0, 0)
h.set_speculation(False)
os.environ.setdefault("G2N_DBG_KEYS", "1000064")  # the table size a warm handle uses on C2 (L2-resident); "" = cold estimate
if not os.environ["G2N_DBG_KEYS"]:
    del os.environ["G2N_DBG_KEYS"]
os.environ["G2N_DBG_TOKENIZE_ONLY"] = "1"
rc = h.build(t.data_ptr(), t.numel(), p)
print("rc", rc)
