"""Per-phase clock stamps of one CTA of the tcgen05 attention kernel (debug build: profiles/experiments/att_trace.sh)."""
import ctypes, os, sys
import numpy as np
import torch
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = ctypes.CDLL(os.path.join(root, "gpurun_libtrace.so"))
lib.pnp_attention_fp16x3_workspace_bytes.restype = ctypes.c_size_t
lib.pnp_attention_fp16x3_workspace_bytes.argtypes = [ctypes.c_int] * 4
lib.pnp_attention_fp16x3.argtypes = [ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                                     ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
dev = torch.device("cuda:0")
B, L, H = 35, 442, 16
qkv = torch.randn(B, L, 3, H, 64, device=dev)
out = torch.empty(B, L, H * 64, device=dev)
n = lib.pnp_attention_fp16x3_workspace_bytes(B, L, H, 64)
ws = torch.empty(n, dtype=torch.uint8, device=dev)
for _ in range(3):
    rc = lib.pnp_attention_fp16x3(qkv.data_ptr(), 1.0, 0.125, out.data_ptr(), None, 1.0, ws.data_ptr(), n, None, B, L, H, 64, None)
    assert rc == 0, rc
torch.cuda.synchronize()
tr = np.zeros((2, 16, 12), dtype=np.int64)
assert lib.pnp_debug_attention_trace(tr.ctypes.data_as(ctypes.c_void_p)) == 0
names = ["top", "S ready", "S read", "pair xch", "S issued/arrived", "O ready", "O folded+V issued", "P written", "arrived p_full", "tile end", "(t0) s_free complete",
         "(t0) p_full complete"]
for who, label in ((0, "thread 0 (issuer)"), (1, "thread 64")):
    print(label)
    t0 = tr[who, 0, 0]
    for j in range(7):
        print("  tile %d: " % j + "  ".join("%s %d" % (names[k], tr[who, j, k] - t0) for k in (0, 1, 2, 3, 10, 4, 5, 6, 7, 8, 11, 9) if tr[who, j, k]))
