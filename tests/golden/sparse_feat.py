"""Plain torch heads shared by tests/golden/make_golden.py (fixture of the real v10Detect3d.inference_forward_feat) and
the GPU test that rebuilds them: imports nothing of the reference."""
import torch


def sparse_feat_modules(C, mid, out_ch, nl, seed):
    """heads[0][i]: the class head on the full map; heads[j][i]: a 5 x 5 conv (no padding: it sees one patch) + 1 x 1 conv
    -- the shapes of v10Detect3d's heads (head.py:560-640).  Seeded on the CPU so that every machine gets the same weights."""
    torch.manual_seed(seed)
    names = list(out_ch)
    heads = [torch.nn.ModuleList(torch.nn.Sequential(torch.nn.Conv2d(C, mid, 3, padding=1), torch.nn.SiLU(),
                                                    torch.nn.Conv2d(mid, out_ch[names[0]], 1)) for _ in range(nl))]
    for n in names[1:]:
        heads.append(torch.nn.ModuleList(torch.nn.Sequential(torch.nn.Conv2d(C, mid, 5, padding=0), torch.nn.SiLU(),
                                                             torch.nn.Conv2d(mid, out_ch[n], 1)) for _ in range(nl)))
    return heads
