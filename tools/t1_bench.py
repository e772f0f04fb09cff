"""T1 alone: table kernel vs tcgen05 kernels, CUDA events on the launch stream, 256 MB L2 flush before every launch.

    python tools/t1_bench.py [N ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaussreg_b200 import ops  # noqa: E402
from gaussreg_b200.config import make_cfg  # noqa: E402
from gaussreg_b200.model import create_model  # noqa: E402


def main():
    ns = [int(a) for a in sys.argv[1:]] or [479, 1024, 4200]
    dev = torch.device("cuda:0")
    model = create_model(make_cfg()).to(dev).eval()
    emb = model.transformer.embedding
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    g = torch.Generator().manual_seed(0)
    for n in ns:
        room = torch.tensor([4.0, 3.0, 2.5]) * (3.0 if n > 2000 else 1.0)  # config 5's room for the large counts
        pts = ((torch.rand(n, 3, generator=g) - 0.5) * room).to(dev)
        d_idx, a_idx, _ = ops.embedding_indices(pts, 0.2, 15, 3)
        res = {}
        outs = {}
        for mode in ("table", "tc"):
            def run():
                if mode == "table":
                    return ops.structure_embedding_tabulated(d_idx, a_idx, emb.embedding.div_term, emb.proj_d.weight, emb.proj_d.bias,
                                                             emb.proj_a.weight, emb.proj_a.bias, 15)
                return ops.structure_embedding_fused(d_idx, a_idx, emb.embedding.div_term, emb.proj_d.weight, emb.proj_d.bias,
                                                     emb.proj_a.weight, emb.proj_a.bias)
            for _ in range(3):
                out = run()
            ts = []
            for _ in range(10):
                del out
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                out = run()
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ts.sort()
            res[mode] = ts[len(ts) // 2]
            outs[mode] = out
        diff = float((outs["table"] - outs["tc"]).norm() / outs["tc"].norm())
        wr = n * n * 256 * 4 / 1e9
        print(f"N={n}: table {res['table']:.3f} ms ({wr / res['table'] * 1e3:.0f} GB/s of output), tc {res['tc']:.3f} ms, "
              f"rel diff {diff:.2e}", flush=True)
        del outs, out


if __name__ == "__main__":
    main()
