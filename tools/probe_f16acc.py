import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pointcloud_rl_b200._lib import lib, stream_ptr
L = lib()
for cfmt, abfmt, name in [(1, 0, "D=f32, A/B=f16"), (0, 0, "D=f16, A/B=f16")]:
    out = torch.zeros(128 * 32, dtype=torch.int32, device="cuda")
    rc = L.cdll.pcrl_debug_f16acc_probe(ctypes.c_int(cfmt), ctypes.c_int(abfmt), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream_ptr()))
    try:
        torch.cuda.synchronize()
    except Exception as e:
        print(name, "-> CUDA error:", str(e)[:200]); break
    w = out.cpu().numpy().reshape(128, 32)
    for r in (0, 37, 100):
        row = w[r]
        if cfmt == 1:
            print(name, f"row {r}: f32 cols 0..7:", row[:8].view(np.float32).tolist())
        else:
            h = row.view(np.float16)
            print(name, f"row {r}: as f16, halfs 0..15:", h[:16].tolist(), "| halfs 16..23:", h[16:24].tolist(), "| halfs 32..39:", h[32:40].tolist(), "| 60..63", h[60:64].tolist())
