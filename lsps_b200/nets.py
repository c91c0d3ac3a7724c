"""Network classes constructible the way the reference constructs its own: `Name(hyperparameters['gen'])`.

The reference trainer builds its nets by name from the YAML (src/trainers/lsps_trainer.py:21-24:
`exec('self.gen = %s(hyperparameters["gen"])' % hyperparameters['gen']['name'])`).  LSPSTrainerB200 shares one Ops /
ParamStore set between its nets and wires them itself; the classes below are the same engines as STAND-ALONE objects --
own parameter store, reference constructor signature, reference `forward` return values (NCHW fp32 tensors) and
state_dict keys -- for code that instantiates or calls a net directly (evaluation scripts, `trainer.gen(x_a, x_b)`).
"""
import torch

from .engine import Ops, Generator, Discriminator, PoseVAE, Mapping
from .params import ParamStore, gen_entries, dis_entries, vae_entries, map_entries


def _device(device):
    if device is None:
        device = torch.cuda.current_device()
    return torch.device("cuda", device if isinstance(device, int) else torch.device(device).index)


def _img(t, dev):
    t = t.detach()
    return t.reshape(t.shape[0], t.shape[-2], t.shape[-1]).to(device=dev, dtype=torch.float32).contiguous()


def _nchw(t):
    """bf16 NHWC -> fp32 tensor with the reference's NCHW shape (a channels_last-strided view, no copy of the layout)"""
    return t.permute(0, 3, 1, 2).float()


class SharedResGenB200(Generator):
    """SharedResGen (lsps_nets.py:164-272); with params['name'] == 'SharedResXGen' the ResNeXt variant (:277-387)."""

    def __init__(self, params, device=None, seed=1, lr=1e-4):
        dev = _device(device)
        store = ParamStore(gen_entries(params), dev, lr, 1e-4)
        store.init_(seed)
        super().__init__(Ops(dev), store, params)
        self._seed, self._draws = seed, 0
        self._kl = torch.zeros(8, device=dev)
        self.state_dict, self.load_state_dict = store.state_dict, store.load_state_dict

    def _noise(self):
        if not self.training:
            return None
        self._draws += 2
        return ("philox", 0x5EED + self._seed, self._draws)

    def forward(self, x_A, x_B):
        """-> (x_Aa, x_Ba, x_Ab, x_Bb, shared) like lsps_nets.py:250-258"""
        xa, xb = _img(x_A, self.ops.device), _img(x_B, self.ops.device)
        oa, ob, z = Generator.forward(self, xa, xb, self._noise(), self._kl)
        n = xa.shape[0]
        u = lambda t: t.unsqueeze(1)
        return u(oa[:n]), u(oa[n:]), u(ob[:n]), u(ob[n:]), _nchw(z)

    __call__ = forward

    def forward_a2b(self, x_A):
        xa = _img(x_A, self.ops.device)
        _, ob, z = Generator.forward(self, xa, None, self._noise(), self._kl)
        return ob.unsqueeze(1), _nchw(z)

    def forward_b2a(self, x_B):
        xb = _img(x_B, self.ops.device)
        oa, _, z = Generator.forward(self, None, xb, self._noise(), self._kl)
        return oa.unsqueeze(1), _nchw(z)

    def decode(self, z):
        """(2n,256,32,32) latent -> (decode_A, decode_B) of ALL latents (lsps_nets.py:239-243)"""
        zz = z.detach().to(self.ops.device).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        y = zz
        for i in range(self.p["n_gen_shared_blk"]):
            y = self.ops.res_fwd(self.S, "dec_shared.%d" % i, y, None)
        return self.dec_fwd("A", y, None).unsqueeze(1), self.dec_fwd("B", y, None).unsqueeze(1)

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)


SharedResXGenB200 = SharedResGenB200     # the block type follows params['name'] / the `.model.6` keys of the store


class SharedDisB200(Discriminator):
    """SharedDis (lsps_nets.py:86-160); split=True runs it on the bf16x3 kernels (the trainer's "mixed" precision)."""

    def __init__(self, params, device=None, seed=2, lr=1e-4, split=True):
        dev = _device(device)
        store = ParamStore(dis_entries(params), dev, lr, 1e-4, split=split)
        store.init_(seed)
        super().__init__(Ops(dev), store, params)

    def forward(self, x_A, x_B):
        """-> (out_A.view(-1), out_B.view(-1), feats_A, feats_B) like lsps_nets.py:154-160"""
        xa, xb = _img(x_A, self.ops.device), _img(x_B, self.ops.device)
        F = self.features(xa, xb)
        lg = self.logits(F)
        na = xa.shape[0]
        c = self.cf
        Ff = (F[..., :c].float() + F[..., c:].float()) if self.split else F.float()
        Ff = Ff.permute(0, 3, 1, 2)
        return lg[:4 * na], lg[4 * na:], Ff[:na], Ff[na:]

    __call__ = forward

    def feats(self, x_aa, x_ba, x_ab, x_bb):
        dev = self.ops.device
        F = self.features(self.ops.cat([_img(x_aa, dev), _img(x_ba, dev)]), self.ops.cat([_img(x_ab, dev), _img(x_bb, dev)]))
        c = self.cf
        Ff = ((F[..., :c].float() + F[..., c:].float()) if self.split else F.float()).permute(0, 3, 1, 2)
        return torch.split(Ff, Ff.shape[0] // 4, 0)


class poseVAEB200(PoseVAE):
    """poseVAE (lsps_nets.py:34-83)"""

    def __init__(self, params, device=None, seed=3, lr=1e-3):
        dev = _device(device)
        store = ParamStore(vae_entries(params), dev, lr, 1e-3)
        store.init_(seed)
        super().__init__(Ops(dev), store, params, lambda shape: torch.randn(shape, device=dev) * 0.05)

    def __call__(self, y):
        return self.forward(y.to(self.ops.device))


class MappingB200(Mapping):
    """Mapping (lsps_nets.py:8-31)"""

    def __init__(self, params, device=None, seed=4, lr=1e-4):
        dev = _device(device)
        store = ParamStore(map_entries(params), dev, lr, 1e-4)
        store.init_(seed)
        super().__init__(Ops(dev), store, params)
