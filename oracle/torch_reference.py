"""Plain-PyTorch restatement of the reference network, for timing the *reference's own code path* (stock
nn.Linear / ATen launches) on whatever device torch offers, and as a second checker.  TEST / BENCH
INFRASTRUCTURE ONLY (same rules as r2l_oracle.py).  Architecture: model/nerf_raybased.py:443-465,:480-544."""
import torch
import torch.nn as nn

from . import r2l_oracle as orc


class RefResMLP(nn.Module):
    def __init__(self, width):
        super().__init__()
        self.body = nn.Sequential(nn.Linear(width, width), nn.ReLU(True), nn.Linear(width, width))

    def forward(self, x):
        return self.body(x).mul(1.0) + x


class RefR2L(nn.Module):
    def __init__(self):
        super().__init__()
        self.head = nn.Sequential(nn.Linear(orc.IN_DIM, orc.WIDTH), nn.ReLU(True))
        self.body = nn.Sequential(*[RefResMLP(orc.WIDTH) for _ in range(orc.N_BLOCKS)])
        self.tail = nn.Sequential(nn.Linear(orc.WIDTH, 3), nn.Sigmoid())

    def forward(self, x):
        x = self.head(x)
        x = self.body(x) + x
        return self.tail(x)

    def load_flat(self, flat: torch.Tensor):
        off = 0
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
        assert off == orc.NUM_PARAMS
        return self

    def flat_grads(self) -> torch.Tensor:
        return torch.cat([p.grad.reshape(-1) for p in self.parameters()])


def embed(pts: torch.Tensor, L: int = orc.N_FREQS) -> torch.Tensor:
    w = 2 ** torch.linspace(0, L - 1, steps=L, device=pts.device)
    y = pts[..., None] * w
    y = torch.cat([torch.sin(y), torch.cos(y), pts.unsqueeze(-1)], dim=-1)
    return y.view(y.shape[0], -1)


def sample(rays_o, rays_d, z_vals):
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[None, :, None]
    return pts.view(pts.shape[0], -1)


# ---- teacher NeRF parameters in the reference's construction order (model/nerf_raybased.py:357-375); test infrastructure ----
def teacher_layer_shapes(D=8, W=256, input_ch=63, input_ch_views=27, skips=(4,)):
    """(out, in) per nn.Linear in the reference's construction order (:357-375)."""
    shapes = [(W, input_ch)] + [(W, W + input_ch) if i in skips else (W, W) for i in range(D - 1)]
    shapes += [(W // 2, input_ch_views + W)]            # views_linears.0
    shapes += [(W, W), (1, W), (3, W // 2)]              # feature_linear, alpha_linear, rgb_linear
    return shapes


def init_teacher_params(seed=None, **kw):
    """Default-initialised parameters drawn in the reference's order; returns [w0, b0, w1, b1, ...] in
    state_dict order (pts_linears.*, views_linears.0, feature_linear, alpha_linear, rgb_linear)."""
    if seed is not None:
        torch.manual_seed(seed)
    out = []
    for o, i in teacher_layer_shapes(**kw):
        lin = nn.Linear(i, o)
        out += [lin.weight.detach().clone(), lin.bias.detach().clone()]
    return out
