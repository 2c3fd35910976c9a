"""CPU restatement of the DCLL hot path (numpy + torch-CPU, float32).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Every function cites the
reference lines it restates (paths relative to /root/reference).

Two training back-ends are provided and are checked against each other and
against the live reference:

  * ``closed``   -- explicit formulas for the local gradient and the Adam update
                    (SURVEY.md section 8a, row a7/a8).  This is what the CUDA
                    kernels implement, stage by stage, so every intermediate
                    (g_o, g_u, gW, gb, gWout ...) can be compared.
  * ``autograd`` -- the same operator sequence the reference runs (F.conv2d,
                    max_pool2d, linear, SmoothL1Loss, ``backward()``,
                    ``torch.optim.Adam.step()``); used as the CPU "port"
                    baseline in bench.py because its cost profile is the
                    reference's.
"""
import math
from collections import Counter
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

GAMMA_EXPONENT_F32 = float(np.float32(1.0 / 1.2))  # torch CPU pow casts the exponent to f32


# ----------------------------------------------------------------------------------------------
# image -> frozen Poisson spike train  (data/utils.py:15-40)
# ----------------------------------------------------------------------------------------------
def image2spiketrain(x, y, input_shape, gain=50, min_duration=None, max_duration=500):
    """data/utils.py:15-40.  x: float32 images [B, ...] (torch tensor or ndarray), y: [B, K].  numpy RNG consumption as the
    reference: one randint for the durations (:26), then per sample one uniform(size=(T_i, Nin)) (:32).  p is float32
    (gain * float32 tensor -> float32; (1000.0 - p32) / 1000 stays float32), the comparison promotes it to float64."""
    if min_duration is None:
        min_duration = max_duration - 1                                       # :18-19
    xa = np.asarray(x, dtype=np.float32)
    batch_size = xa.shape[0]                                                  # :21
    nin = int(np.prod(input_shape))                                           # :22
    rates = (np.float32(gain) * xa.reshape(batch_size, -1)).astype(np.float32)    # :23
    p = ((np.float32(1000.0) - rates) / np.float32(1000)).astype(np.float32)  # :24
    T = np.random.randint(min_duration, max_duration, batch_size)            # :25
    all_inputs = np.zeros((max_duration, batch_size, nin))                    # :28
    for i in range(batch_size):                                               # :29-32
        spikes = np.ones((T[i], nin))
        spikes[(np.random.uniform(size=(T[i], nin)) < p[i]).astype('bool')] = 0
        all_inputs[:T[i], i, :] = spikes
    all_inputs = all_inputs.reshape(max_duration, batch_size, *input_shape)   # :35
    all_target = np.repeat(np.asarray(y)[np.newaxis, :, :], max_duration, axis=0)   # :38
    return all_inputs, all_target


# ----------------------------------------------------------------------------------------------
# IQ -> spike encoding  (data/utils.py:43-87)
# ----------------------------------------------------------------------------------------------
def encode_cells(x, out_w=28, out_h=28, min_I=-1, max_I=1, min_Q=-1, max_Q=1, t_start=0,
                 max_duration=500, do_gamma=True):
    """Cell indices of iq2spiketrain as int32 [T, B, 2] = (cell_Q (row), cell_I (col)).

    data/utils.py:60-79.  ``x``: float32 (B, 2, N) (or (B,2,1,N), squeezed as at :47).
    float32 arithmetic, one rounding per reference operation; the gamma power is
    evaluated in double with the float32-rounded exponent and rounded once to
    float32, which reproduces torch-CPU's ``pow`` cell-for-cell.
    """
    x = np.asarray(x, dtype=np.float32)
    x = x.reshape(x.shape[0], 2, x.shape[-1])
    f32 = np.float32

    def axis(v, lo, hi, n):
        c = (v - f32(lo)) / f32(hi - lo)                      # :65-66
        if do_gamma:
            c = c * f32(2.0) - f32(1.0)                       # :69-70
            mag = np.power(np.abs(c).astype(np.float64), GAMMA_EXPONENT_F32).astype(np.float32)
            c = np.sign(c) * mag                              # :72-73
            c = (c + f32(1.0)) * f32(0.5)                     # :75-76
        c = np.clip(c, f32(0.0), f32(1.0)) * f32(n - 1)       # :78-79
        return c.astype(np.int32)                             # .int(): truncation

    seg = x[:, :, t_start:t_start + max_duration]
    cell_I = axis(seg[:, 0, :], min_I, max_I, out_w)          # (B, T)
    cell_Q = axis(seg[:, 1, :], min_Q, max_Q, out_h)
    return np.ascontiguousarray(np.stack([cell_Q.T, cell_I.T], axis=-1))  # (T, B, 2)


def cells_to_frames(cells, out_h, out_w, dtype=np.float32):
    """Dense one-hot frames [T, B, 1, H, W] (data/utils.py:57,81-82): spike at [cell_Q, cell_I]."""
    T, B, _ = cells.shape
    frames = np.zeros((T, B, 1, out_h, out_w), dtype=dtype)
    tt, bb = np.meshgrid(np.arange(T), np.arange(B), indexing="ij")
    frames[tt, bb, 0, cells[..., 0], cells[..., 1]] = 1
    return frames


def to_one_hot(t, width):
    """data/utils.py:10-12."""
    t = torch.as_tensor(t, dtype=torch.int64)
    out = torch.zeros(*t.shape, width)
    return out.scatter_(1, t.unsqueeze(-1), 1)


def repeat_targets(y, T):
    """data/utils.py:85."""
    return np.repeat(np.asarray(y)[np.newaxis, :, :], T, axis=0)


# ----------------------------------------------------------------------------------------------
# Layer containers
# ----------------------------------------------------------------------------------------------
def _pair(v):
    return (int(v), int(v)) if not hasattr(v, "__len__") else (int(v[0]), int(v[1]))


@dataclass
class ConvSpec:
    in_channels: int
    out_channels: int
    kernel_size: Tuple[int, int]
    padding: Tuple[int, int]
    pooling: Tuple[int, int]
    im_dims: Tuple[int, int]
    target_size: int
    alpharp: float = 0.65
    wrp: float = 0.0
    output_layer: bool = False

    @property
    def conv_shape(self):  # dcll/pytorch_libdcll.py:368-375 (stride 1, dilation 1)
        return (self.im_dims[0] + 2 * self.padding[0] - self.kernel_size[0] + 1,
                self.im_dims[1] + 2 * self.padding[1] - self.kernel_size[1] + 1)

    @property
    def pool_pad(self):  # :545-546
        return ((self.pooling[0] - 1) // 2, (self.pooling[1] - 1) // 2)

    @property
    def out_shape(self):  # :593-597
        h, w = self.conv_shape
        return (h // self.pooling[0], w // self.pooling[1])

    @property
    def flat_size(self):
        h, w = self.out_shape
        return self.out_channels * h * w


@dataclass
class ConvParams:
    weight: torch.Tensor
    bias: torch.Tensor
    alpha: torch.Tensor
    alphas: torch.Tensor
    tau_m: torch.Tensor
    tau_s: torch.Tensor
    wo: torch.Tensor
    bo: torch.Tensor
    wout: Optional[torch.Tensor] = None
    bout: Optional[torch.Tensor] = None


@dataclass
class ConvState:
    eps0: torch.Tensor
    eps1: torch.Tensor
    arp: Optional[torch.Tensor] = None


@dataclass
class AdamSlot:
    step: int = 0
    exp_avg: Optional[torch.Tensor] = None
    exp_avg_sq: Optional[torch.Tensor] = None


def zero_state(spec: ConvSpec, batch):
    """dcll/pytorch_libdcll.py:377-383 / :468-477."""
    shp = (batch, spec.in_channels) + tuple(spec.im_dims)
    arp = torch.zeros((batch, spec.out_channels) + spec.conv_shape) if spec.wrp > 0 else None
    return ConvState(torch.zeros(shp), torch.zeros(shp), arp)


# ----------------------------------------------------------------------------------------------
# Layer forward  (ContinuousConv2D.forward :407-426, RRP :485-509, Conv2dDCLLlayer.forward :599-608)
# ----------------------------------------------------------------------------------------------
@dataclass
class FwdOut:
    output: torch.Tensor          # pooled spikes, or logits of output_ on the last layer (:605-606)
    spikes: torch.Tensor          # pooled spikes (always)
    pvoutput: torch.Tensor        # [B, K] local readout (:602-603)
    pv: torch.Tensor              # pooled sigmoid [B, Cout, Hp, Wp]
    pvmem: torch.Tensor           # membrane (incl. refractory term for RRP) [B, Cout, H', W']
    pool_idx: torch.Tensor        # flat argmax index into H'*W' of every pooled pv cell
    state: ConvState = None


def conv_step_fwd(spec: ConvSpec, p: ConvParams, st: ConvState, x: torch.Tensor, weight=None) -> FwdOut:
    w = p.weight if weight is None else weight
    eps0 = x * p.tau_s + p.alphas * st.eps0                                   # :415
    eps1 = p.alpha * st.eps1 + eps0 * p.tau_m                                 # :416
    pvmem = F.conv2d(eps1, w, p.bias, 1, spec.padding, 1, 1)                  # :417-418
    if spec.wrp > 0:
        arp = spec.alpharp * st.arp                                           # :497
        u = pvmem + arp                                                       # :498
        spikes = (u > 0).float()                                              # :499
        pv = torch.sigmoid(u)                                                 # :500
        arp = arp - spikes * spec.wrp                                         # :503
    else:
        u, arp = pvmem, None
        pv = torch.sigmoid(u)                                                 # :419
        spikes = (u > 0).float()                                              # :420
    k, pad = spec.pooling, spec.pool_pad
    sp_p = F.max_pool2d(spikes, k, k, pad)                                    # :601
    pv_p, idx = F.max_pool2d(pv, k, k, pad, return_indices=True)              # :601
    flat = pv_p.reshape(pv_p.shape[0], -1)                                    # :602
    pvoutput = F.linear(flat, p.wo, p.bo)                                     # :603
    output = sp_p
    if spec.output_layer:
        output = F.linear(flat.detach(), p.wout, p.bout)                      # :605-606
    new_state = ConvState(eps0.detach(), eps1.detach(), None if arp is None else arp.detach())
    return FwdOut(output, sp_p, pvoutput, pv_p, u, idx, new_state)


# ----------------------------------------------------------------------------------------------
# Local loss gradient, closed form  (DCLLBase.train_dcll :690-704 + autograd)
# ----------------------------------------------------------------------------------------------
def loss_grad(pred, target, kind="smoothl1"):
    """d(mean-reduced loss)/d(pred) for the loss classes train.py:173 can select."""
    d = pred - target
    n = d.numel()
    if kind == "smoothl1":     # torch.nn.SmoothL1Loss(beta=1): 0.5 d^2 if |d|<1 else |d|-0.5
        g = torch.where(d.abs() < 1, d, torch.sign(d))
    elif kind == "mse":
        g = 2 * d
    elif kind == "l1":
        g = torch.sign(d)
    else:
        raise ValueError(kind)
    return g / n


def loss_value(pred, target, kind="smoothl1"):
    if kind == "smoothl1":
        return F.smooth_l1_loss(pred, target)
    if kind == "mse":
        return F.mse_loss(pred, target)
    if kind == "l1":
        return F.l1_loss(pred, target)
    raise ValueError(kind)


@dataclass
class LocalGrads:
    g_o: torch.Tensor
    g_u: torch.Tensor            # dL/d(membrane) [B, Cout, H', W'] (zero where not the pool argmax)
    gW: torch.Tensor
    gb: torch.Tensor
    g_o2: Optional[torch.Tensor] = None
    gWout: Optional[torch.Tensor] = None
    gbout: Optional[torch.Tensor] = None


def conv_local_grads(spec: ConvSpec, p: ConvParams, fo: FwdOut, target, loss_kind="smoothl1") -> LocalGrads:
    B = fo.pv.shape[0]
    g_o = loss_grad(fo.pvoutput, target, loss_kind)                           # :694
    g_flat = g_o @ p.wo                                                       # d pvoutput / d flat
    g_pv = g_flat.reshape(fo.pv.shape)
    g_pool = g_pv * fo.pv * (1 - fo.pv)                                       # sigmoid' at the argmax
    Hc, Wc = spec.conv_shape
    g_u = torch.zeros(B, spec.out_channels, Hc * Wc)
    g_u.scatter_(2, fo.pool_idx.reshape(B, spec.out_channels, -1), g_pool.reshape(B, spec.out_channels, -1))
    g_u = g_u.reshape(B, spec.out_channels, Hc, Wc)
    eps1 = fo.state.eps1
    gW = torch.nn.grad.conv2d_weight(eps1, p.weight.shape, g_u, 1, spec.padding, 1, 1)
    gb = g_u.sum(dim=(0, 2, 3))
    out = LocalGrads(g_o, g_u, gW, gb)
    if spec.output_layer:
        out.g_o2 = loss_grad(fo.output, target, loss_kind)                    # :696
        flat = fo.pv.reshape(B, -1)
        out.gWout = out.g_o2.t() @ flat
        out.gbout = out.g_o2.sum(0)
    return out


def adam_update(w, g, slot: AdamSlot, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """torch.optim.Adam single-tensor step (torch/optim/adam.py, _single_tensor_adam), float32.

    Reference call sites: dcll/pytorch_libdcll.py:634-638,711-714; hyper-parameters
    train.py:164-168 (betas [0, beta], weight_decay 10).  In place on ``w``.
    """
    if slot.exp_avg is None:
        slot.exp_avg = torch.zeros_like(w)
        slot.exp_avg_sq = torch.zeros_like(w)
    slot.step += 1
    if weight_decay != 0:
        g = g.add(w, alpha=weight_decay)
    slot.exp_avg.lerp_(g, 1 - beta1)
    slot.exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** slot.step
    bc2 = 1 - beta2 ** slot.step
    step_size = lr / bc1
    denom = (slot.exp_avg_sq.sqrt() / math.sqrt(bc2)).add_(eps)
    w.addcdiv_(slot.exp_avg, denom, value=-step_size)
    return w


# ----------------------------------------------------------------------------------------------
# Vote accuracy  (dcll/pytorch_libdcll.py:44-61, :735-749)
# ----------------------------------------------------------------------------------------------
def predictions_by_vote(clout, labels_onehot):
    """clout: sequence of [B] int arrays (one per counted timestep); labels_onehot [T', B, K]."""
    pv = np.array(clout).T
    lab = np.asarray(labels_onehot).argmax(axis=2).T
    n = len(pv)
    pred = np.empty(n)
    true = np.empty(n)
    for i in range(n):
        pred[i] = Counter(pv[i].tolist()).most_common(1)[0][0]
        true[i] = Counter(lab[i].tolist()).most_common(1)[0][0]
    return pred, true


def accuracy_by_vote(clout, labels_onehot):
    labels_onehot = np.asarray(labels_onehot)
    begin = len(clout)
    pred, true = predictions_by_vote(clout, labels_onehot[-begin:])           # :736-737
    return float(np.mean(pred == true))


def confusion_matrix(clout, labels_onehot, num_classes):
    labels_onehot = np.asarray(labels_onehot)
    pred, true = predictions_by_vote(clout, labels_onehot[-len(clout):])
    cm = np.zeros((num_classes, num_classes), dtype=int)                      # :745-748
    for a, b in zip(pred, true):
        cm[int(a), int(b)] += 1
    return cm


# ----------------------------------------------------------------------------------------------
# Network  (networks/__init__.py:116-199 + DCLLBase/DCLLClassification :615-749)
# ----------------------------------------------------------------------------------------------
BUILTIN_SPECS = {
    # values of networks/radio_ml_conv.yaml, mnist_conv.yaml, radio_ml_conv_ref.yaml
    "radio_ml_conv": [dict(out_channels=32, kernel_size=7, padding=3, pooling=1)] * 3,
    "mnist_conv": [dict(out_channels=16, kernel_size=7, padding=2, pooling=2),
                   dict(out_channels=24, kernel_size=7, padding=2, pooling=1),
                   dict(out_channels=32, kernel_size=7, padding=2, pooling=2)],
    "radio_ml_conv_ref": [dict(out_channels=64, kernel_size=(1, 3), padding=(0, 1), pooling=(1, 2))] * 7,
}


def make_specs(convs, im_dims, target_size, alpharp=0.65, wrp=0.0, netscale=1.0):
    """Chain of ConvSpec as ConvNetwork.__init__ builds it (networks/__init__.py:126-151)."""
    specs = []
    cin, hw = im_dims[0], tuple(im_dims[1:3])
    for i, c in enumerate(convs):
        s = ConvSpec(cin, int(c["out_channels"] * netscale), _pair(c["kernel_size"]), _pair(c["padding"]),
                     _pair(c["pooling"]), hw, target_size, alpharp, wrp, output_layer=(i == len(convs) - 1))
        specs.append(s)
        cin, hw = s.out_channels, s.out_shape
    return specs


def params_from_state_dict(sd, n_layers):
    """Oracle parameters from a reference ConvNetwork.state_dict() (key names SURVEY section 5)."""
    out = []
    for i in range(n_layers):
        pre = "dcll_slices.%d.dclllayer." % i
        g = lambda k: sd[pre + k].detach().clone().float()
        p = ConvParams(g("i2h.weight"), g("i2h.bias"), g("i2h.alpha"), g("i2h.alphas"), g("i2h.tau_m__dt"),
                       g("i2h.tau_s__dt"), g("i2o.weight"), g("i2o.bias"))
        if pre + "output_.weight" in sd:
            p.wout, p.bout = g("output_.weight"), g("output_.bias")
        out.append(p)
    return out


class OracleNet:
    """ConvNetwork + DCLLClassification slices, functional restatement.

    ``learn``/``test``/``reset``/``accuracy`` follow networks/__init__.py:175-199;
    the per-slice step follows dcll/pytorch_libdcll.py:690-729.  Neuron state is
    NOT cleared by ``reset()``.
    """

    def __init__(self, specs: List[ConvSpec], params: List[ConvParams], batch_size, burnin=50,
                 lrs=(1e-6,), betas=(0.0, 0.95), weight_decay=10.0, adam_eps=1e-8, lr_out=1e-4,
                 loss_kind="smoothl1", backend="closed"):
        self.specs, self.params = specs, params
        self.batch_size, self.burnin = batch_size, burnin
        self.lrs = [lrs[min(i, len(lrs) - 1)] for i in range(len(specs))]   # networks/__init__.py:154-157
        self.betas, self.weight_decay, self.adam_eps, self.lr_out = betas, weight_decay, adam_eps, lr_out
        self.loss_kind, self.backend = loss_kind, backend
        self.states = [zero_state(s, batch_size) for s in specs]
        self.slots = [dict(w=AdamSlot(), b=AdamSlot(), wout=AdamSlot(), bout=AdamSlot()) for _ in specs]
        self.iters = [0] * len(specs)
        self.clout = [[] for _ in specs]
        self.last = [None] * len(specs)   # FwdOut of the latest step (for parity probes)
        self.last_grads = [None] * len(specs)
        self._opts = None

    # -- networks/__init__.py:187-189 -> DCLLBase.init :648-653
    def reset(self, init_states=False):
        for i in range(len(self.specs)):
            self.clout[i] = []
            self.iters[i] = 0
            if init_states:
                self.states[i] = zero_state(self.specs[i], self.batch_size)

    def _record(self, i, fo, ignore_burnin):
        if ignore_burnin or self.iters[i] >= self.burnin:                     # :724
            src = fo.output if self.specs[i].output_layer else fo.pvoutput    # :725-728
            self.clout[i].append(src.argmax(1).numpy())

    def test(self, x):
        spikes = x
        for i, (s, p) in enumerate(zip(self.specs, self.params)):
            self.iters[i] += 1                                                # :656
            with torch.no_grad():
                fo = conv_step_fwd(s, p, self.states[i], spikes)
            self.states[i] = fo.state
            self.last[i] = fo
            self._record(i, fo, True)
            spikes = fo.output

    def learn(self, x, target):
        if self.backend == "autograd":
            return self._learn_autograd(x, target)
        spikes = x
        for i, (s, p) in enumerate(zip(self.specs, self.params)):
            self.iters[i] += 1
            with torch.no_grad():
                fo = conv_step_fwd(s, p, self.states[i], spikes)
                self.states[i] = fo.state
                self.last[i] = fo
                self._record(i, fo, False)
                if self.iters[i] >= self.burnin:                              # :692
                    g = conv_local_grads(s, p, fo, target, self.loss_kind)
                    self.last_grads[i] = g
                    sl = self.slots[i]
                    kw = dict(lr=self.lrs[i], beta1=self.betas[0], beta2=self.betas[1], eps=self.adam_eps,
                              weight_decay=self.weight_decay)
                    adam_update(p.weight, g.gW, sl["w"], **kw)                # :712
                    adam_update(p.bias, g.gb, sl["b"], **kw)
                    if s.output_layer:                                        # :713-714, optimizer2 lr 1e-4 defaults
                        adam_update(p.wout, g.gWout, sl["wout"], lr=self.lr_out)
                        adam_update(p.bout, g.gbout, sl["bout"], lr=self.lr_out)
            spikes = fo.output

    # same operator sequence as the reference (autograd + torch.optim.Adam)
    def _learn_autograd(self, x, target):
        if self._opts is None:
            self._opts = []
            for i, p in enumerate(self.params):
                p.weight.requires_grad_(True)
                p.bias.requires_grad_(True)
                o1 = torch.optim.Adam([p.weight, p.bias], lr=self.lrs[i], betas=list(self.betas),
                                      weight_decay=self.weight_decay, eps=self.adam_eps)
                o2 = None
                if self.specs[i].output_layer:
                    p.wout.requires_grad_(True)
                    p.bout.requires_grad_(True)
                    o2 = torch.optim.Adam([p.wout, p.bout], lr=self.lr_out)
                self._opts.append((o1, o2))
        spikes = x
        for i, (s, p) in enumerate(zip(self.specs, self.params)):
            self.iters[i] += 1
            fo = conv_step_fwd(s, p, self.states[i], spikes)
            self.states[i] = fo.state
            self.last[i] = fo
            self._record(i, fo, False)
            if self.iters[i] >= self.burnin:
                o1, o2 = self._opts[i]
                o1.zero_grad()
                if o2 is not None:
                    o2.zero_grad()
                loss = loss_value(fo.pvoutput, target, self.loss_kind)
                if s.output_layer:
                    loss = loss + loss_value(fo.output, target, self.loss_kind)
                loss.backward()
                o1.step()
                if o2 is not None:
                    o2.step()
            spikes = fo.output.detach()

    def accuracy(self, labels_onehot):
        return [accuracy_by_vote(c, labels_onehot) for c in self.clout]


def random_params(specs, seed=0, random_tau=True, lc_ampl=0.5):
    """Parameters drawn with the reference's distributions (not its RNG stream).

    reset_parameters :359-366, reset_lc_parameters :583-587, randomize_tau :391-405.
    Used for oracle-vs-CUDA tests that do not need the reference's exact draw.
    """
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    out = []
    for s in specs:
        n = s.in_channels * s.kernel_size[0] * s.kernel_size[1]
        stdv = 1.0 / math.sqrt(n) / 250
        u = lambda shape, a: (torch.rand(shape, generator=g) * 2 - 1) * a
        w = u((s.out_channels, s.in_channels) + tuple(s.kernel_size), stdv * 1e-2)
        b = u((s.out_channels,), stdv)
        if random_tau:
            taum = rs.uniform(5, 35, size=[s.in_channels]) * 1e-3
            taus = rs.uniform(5, 10, size=[s.in_channels]) * 1e-3
            bc = lambda t: np.broadcast_to(t, tuple(s.im_dims) + (s.in_channels,)).transpose(2, 0, 1)
            alpha = torch.Tensor(1 - 1e-3 / bc(taum))
            alphas = torch.Tensor(1 - 1e-3 / bc(taus))
        else:
            alpha, alphas = torch.Tensor([0.92]), torch.Tensor([0.85])
        tau_m, tau_s = 1.0 / (1 - alpha), 1.0 / (1 - alphas)
        F_ = s.flat_size
        so = lc_ampl / math.sqrt(F_)
        p = ConvParams(w, b, alpha, alphas, tau_m, tau_s, u((s.target_size, F_), so), u((s.target_size,), so))
        if s.output_layer:
            k = 1.0 / math.sqrt(F_)
            p.wout, p.bout = u((s.target_size, F_), k), u((s.target_size,), k)
        out.append(p)
    return out


# ----------------------------------------------------------------------------------------------
# Dense layer  (CLLDenseModule.forward :131-148, RRP :171-195, DenseDCLLlayer.forward :250-255)
# ----------------------------------------------------------------------------------------------
@dataclass
class DenseFwdOut:
    spikes: torch.Tensor
    pvoutput: torch.Tensor
    pv: torch.Tensor
    vmem: torch.Tensor
    state: ConvState


def dense_step_fwd(p: ConvParams, st: ConvState, x, alpharp=0.65, wrp=0.0) -> DenseFwdOut:
    x = x.reshape(-1, p.weight.shape[1])                                      # :251
    eps0 = x * p.tau_s + p.alphas * st.eps0                                   # :139
    eps1 = p.alpha * st.eps1 + eps0 * p.tau_m                                 # :140
    vmem = F.linear(eps1, p.weight, p.bias)                                   # :141
    arp = None
    if wrp > 0:
        arp = alpharp * st.arp                                                # :182
        vmem = vmem + arp                                                     # :183
    spikes = (vmem > 0).float()
    pv = torch.sigmoid(vmem)
    if wrp > 0:
        arp = arp - spikes * wrp                                              # :189
    pvoutput = F.linear(pv, p.wo, p.bo)                                       # :253
    return DenseFwdOut(spikes, pvoutput, pv, vmem, ConvState(eps0, eps1, arp))


def dense_local_grads(p: ConvParams, fo: DenseFwdOut, target, loss_kind="smoothl1"):
    """gW = ((g_o Wo) * s(1-s))^T eps1."""
    g_o = loss_grad(fo.pvoutput, target, loss_kind)
    g_u = (g_o @ p.wo) * fo.pv * (1 - fo.pv)
    return g_o, g_u, g_u.t() @ fo.state.eps1, g_u.sum(0)
