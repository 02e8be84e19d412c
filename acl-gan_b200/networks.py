"""Drop-in `networks` module of the B200-native ACL-GAN implementation.

Same public classes, constructor signatures, child-module names and `state_dict` keys as the reference's
networks.py (AdaINGen / MsImageDis / the block classes - SURVEY.md 8b), so checkpoints, `weights_init`
(utils.py:274-294) and the optimizer parameter order interchange.  The modules own the fp32 OIHW master
parameters; all arithmetic is done by `engine.Engine` through the CUDA extension - there is no eager /
CPU forward: calling a network without a CUDA device raises.

Two call levels:
  * reference-compatible tensor API (`encode`, `decode`, `forward`, `calc_*_loss`): NCHW fp32 in / out,
    forward values only (no autograd graph is attached to the results);
  * engine API (`enc(...)`, `dec(...)`, `dis(...)`) used by trainer.py, which records the hand-scheduled
    backward pass on an `engine.Tape`.
"""
import ctypes as C
import os

import torch
import torch.nn.functional as F
from torch import nn

import aclgan_native as N
import engine as E

_ACT = {"relu": N.ACT_RELU, "lrelu": N.ACT_LRELU, "tanh": N.ACT_TANH, "none": N.ACT_NONE}
_GAN_KIND = {"lsgan": N.GAN_LSGAN, "nsgan": N.GAN_NSGAN}     # MsImageDis gan_type (reference networks.py:66-74)
_NORM = {"none": N.NORM_NONE, "in": N.NORM_IN, "adain": N.NORM_ADAIN, "ln": N.NORM_LN}

_default_engine = {}


def get_engine(precision=None):
    """process-wide engine per precision ('bf16' throughput mode, 'fp32x3' parity mode)"""
    import os
    precision = precision or os.environ.get("ACLGAN_PRECISION", "bf16")
    if precision not in _default_engine:
        _default_engine[precision] = E.Engine(precision)
    return _default_engine[precision]


# ======================================================================================================
# normalisation layers (parameter / buffer holders; the arithmetic lives in csrc/elementwise.cu)
# ======================================================================================================
class AdaptiveInstanceNorm2d(nn.Module):
    """reference networks.py:477-506; weight / bias are assigned per forward by AdaINGen.decode"""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = None
        self.bias = None
        self.register_buffer("running_mean", torch.zeros(num_features))   # never updated (networks.py:487-494)
        self.register_buffer("running_var", torch.ones(num_features))

    def __repr__(self):
        return "%s(%d)" % (self.__class__.__name__, self.num_features)


class LayerNorm(nn.Module):
    """reference networks.py:509-536 (per-sample mean / unbiased std, eps outside, per-channel affine)"""

    def __init__(self, num_features, eps=1e-5, affine=True):
        super().__init__()
        self.num_features, self.affine, self.eps = num_features, affine, eps
        if affine:
            self.gamma = nn.Parameter(torch.Tensor(num_features).uniform_())
            self.beta = nn.Parameter(torch.zeros(num_features))


def l2normalize(v, eps=1e-12):
    return v / (v.norm() + eps)


class SpectralNorm(nn.Module):
    """option surface only (norm='sn' is not reachable from any shipped config)"""

    def __init__(self, module, name="weight", power_iterations=1):
        super().__init__()
        raise NotImplementedError("spectral norm ('sn') is outside the B200 hot path")


# ======================================================================================================
# blocks
# ======================================================================================================
class Conv2dBlock(nn.Module):
    """pad -> conv -> norm -> activation (reference networks.py:312-371)"""

    def __init__(self, input_dim, output_dim, kernel_size, stride, padding=0, norm="none", activation="relu",
                 pad_type="zero"):
        super().__init__()
        self.use_bias = True
        if pad_type == "reflect":
            self.pad = nn.ReflectionPad2d(padding)
        elif pad_type in ("replicate", "zero"):
            raise NotImplementedError("pad_type %r: only 'reflect' is implemented on the B200 path" % pad_type)
        else:
            assert 0, "Unsupported padding type: {}".format(pad_type)
        if norm == "in":
            self.norm = nn.InstanceNorm2d(output_dim)
        elif norm == "ln":
            self.norm = LayerNorm(output_dim)
        elif norm == "adain":
            self.norm = AdaptiveInstanceNorm2d(output_dim)
        elif norm == "none":
            self.norm = None
        elif norm in ("bn", "sn"):
            raise NotImplementedError("norm %r is outside the B200 hot path" % norm)
        else:
            assert 0, "Unsupported normalization: {}".format(norm)
        if activation in ("relu", "lrelu", "tanh"):
            self.activation = {"relu": nn.ReLU(inplace=True), "lrelu": nn.LeakyReLU(0.2, inplace=True),
                               "tanh": nn.Tanh()}[activation]
        elif activation == "none":
            self.activation = None
        elif activation in ("prelu", "selu"):
            raise NotImplementedError("activation %r is outside the B200 hot path" % activation)
        else:
            assert 0, "Unsupported activation: {}".format(activation)
        self.conv = nn.Conv2d(input_dim, output_dim, kernel_size, stride, bias=self.use_bias)
        self.spec = dict(cin=input_dim, cout=output_dim, k=kernel_size, stride=stride, pad=padding,
                         norm=norm, act=activation)
        self._layer = None

    def layer(self, eng, arena, window=N.WINDOW_NONE):
        if self._layer is None or self._layer.eng is not eng or self._layer.arena is not arena:
            self._layer = E.ConvLayer(eng, arena, self.conv.weight, self.conv.bias, self.spec["stride"],
                                      self.spec["pad"], window)
        return self._layer

    def forward(self, x):
        raise NotImplementedError("Conv2dBlock is executed by its owning network through the CUDA engine")


class ResBlock(nn.Module):
    def __init__(self, dim, norm="in", activation="relu", pad_type="zero"):
        super().__init__()
        self.model = nn.Sequential(
            Conv2dBlock(dim, dim, 3, 1, 1, norm=norm, activation=activation, pad_type=pad_type),
            Conv2dBlock(dim, dim, 3, 1, 1, norm=norm, activation="none", pad_type=pad_type))


class ResBlocks(nn.Module):
    def __init__(self, num_blocks, dim, norm="in", activation="relu", pad_type="zero"):
        super().__init__()
        self.model = nn.Sequential(*[ResBlock(dim, norm=norm, activation=activation, pad_type=pad_type)
                                     for _ in range(num_blocks)])


class LinearBlock(nn.Module):
    def __init__(self, input_dim, output_dim, norm="none", activation="relu"):
        super().__init__()
        if norm != "none":
            raise NotImplementedError("LinearBlock norm %r is outside the B200 hot path" % norm)
        self.fc = nn.Linear(input_dim, output_dim, bias=True)
        self.norm = None
        if activation == "relu":
            self.activation = nn.ReLU(inplace=True)
        elif activation == "none":
            self.activation = None
        else:
            raise NotImplementedError("LinearBlock activation %r" % activation)


class MLP(nn.Module):
    """style code -> AdaIN parameters (reference networks.py:280-292)"""

    def __init__(self, input_dim, output_dim, dim, n_blk, norm="none", activ="relu"):
        super().__init__()
        blocks = [LinearBlock(input_dim, dim, norm=norm, activation=activ)]
        blocks += [LinearBlock(dim, dim, norm=norm, activation=activ) for _ in range(n_blk - 2)]
        blocks += [LinearBlock(dim, output_dim, norm="none", activation="none")]
        self.model = nn.Sequential(*blocks)


class StyleEncoder(nn.Module):
    def __init__(self, n_downsample, input_dim, dim, style_dim, norm, activ, pad_type):
        super().__init__()
        m = [Conv2dBlock(input_dim, dim, 7, 1, 3, norm=norm, activation=activ, pad_type=pad_type)]
        for _ in range(2):
            m.append(Conv2dBlock(dim, 2 * dim, 4, 2, 1, norm=norm, activation=activ, pad_type=pad_type))
            dim *= 2
        for _ in range(n_downsample - 2):
            m.append(Conv2dBlock(dim, dim, 4, 2, 1, norm=norm, activation=activ, pad_type=pad_type))
        m.append(nn.AdaptiveAvgPool2d(1))
        m.append(nn.Conv2d(dim, style_dim, 1, 1, 0))
        self.model = nn.Sequential(*m)
        self.output_dim = dim


class ContentEncoder(nn.Module):
    def __init__(self, n_downsample, n_res, input_dim, dim, norm, activ, pad_type):
        super().__init__()
        m = [Conv2dBlock(input_dim, dim, 7, 1, 3, norm=norm, activation=activ, pad_type=pad_type)]
        for _ in range(n_downsample):
            m.append(Conv2dBlock(dim, 2 * dim, 4, 2, 1, norm=norm, activation=activ, pad_type=pad_type))
            dim *= 2
        m.append(ResBlocks(n_res, dim, norm=norm, activation=activ, pad_type=pad_type))
        self.model = nn.Sequential(*m)
        self.output_dim = dim


class Decoder(nn.Module):
    def __init__(self, n_upsample, n_res, dim, output_dim, res_norm="adain", activ="relu", pad_type="zero"):
        super().__init__()
        m = [ResBlocks(n_res, dim, res_norm, activ, pad_type=pad_type)]
        for _ in range(n_upsample):
            m += [nn.Upsample(scale_factor=2),
                  Conv2dBlock(dim, dim // 2, 5, 1, 2, norm="ln", activation=activ, pad_type=pad_type)]
            dim //= 2
        m.append(Conv2dBlock(dim, output_dim, 7, 1, 3, norm="none", activation="tanh", pad_type=pad_type))
        self.model = nn.Sequential(*m)


class _EngineNet(nn.Module):
    """shared plumbing: lazily binds the parameters to an engine + gradient arena"""

    def __init__(self):
        super().__init__()
        self._eng = None
        self._arena = None
        self._bound = False
        self._own_arena = False
        self.train_weights = True
        self._igraphs = {}
        self.infer_graphs = os.environ.get("ACLGAN_INFER_GRAPHS", "1") != "0"
        self.register_load_state_dict_post_hook(lambda module, keys: module.mark_dirty())

    def _repack_dirty(self):
        for l in self.conv_layers():
            if l.dirty:
                l.repack()

    def _graphed(self, key, fn, inputs, owners=None):
        """forward-only inference as a cached CUDA graph (reference: the eager encode / decode / sample calls of test.py:96-106
        and trainer.py:179-245, a few hundred launches each at batch 1): static input buffers -> static outputs, cloned for
        the caller.  Weights are read through the packed buffers the optimizer kernel rewrites in place, so a captured
        graph stays valid across training steps; masters replaced by load_state_dict are repacked before the replay."""
        for net in (owners or (self,)):
            net._ensure_bound()
            net._repack_dirty()
        if not self.infer_graphs or not torch.cuda.is_available():
            return fn(*inputs)
        ent = self._igraphs.get(key)
        if ent is None:
            static_in = [torch.empty_like(t) for t in inputs]
            for st, t in zip(static_in, inputs):
                st.copy_(t)
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):           # warm-up off the capture stream (lazy kernel / attribute initialisation)
                fn(*static_in)
            cur.wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = fn(*static_in)
            ent = (graph, static_in, outs)
            self._igraphs[key] = ent
        graph, static_in, outs = ent
        for st, t in zip(static_in, inputs):
            st.copy_(t, non_blocking=True)
        graph.replay()
        return tuple(o.clone() for o in outs)

    def bind(self, eng, arena=None):
        """(re)binds the network to an engine and a gradient arena: layers, gradient slices and cached inference graphs of a
        previous binding (e.g. stand-alone encode / decode before the trainer's first update) are dropped"""
        self._eng = eng
        self._own_arena = arena is None
        self._arena = arena if arena is not None else E.GradArena(eng.device)
        self._bound = False
        self._dense_grads = []
        self._igraphs = {}

    def _ensure_bound(self):
        if self._eng is None:
            self.bind(get_engine(getattr(self, "precision_hint", None)))      # (the owning trainer's `precision` setting)
        if not self._bound:
            for p in self.parameters():
                if p.device.type != "cuda":
                    raise N.NativeError("aclgan_b200: parameters must live on a CUDA device (no CPU path)")
            self._build_layers()
            self._bound = True
            if self._own_arena:
                self._arena.finalize()
                self.attach_grads()

    def conv_layers(self):
        return [m._layer for m in self.modules() if isinstance(m, Conv2dBlock) and m._layer is not None]

    def mark_dirty(self):
        for l in self.conv_layers():
            l.dirty = True

    def derive_weights(self):
        """re-derives every weight set computed FROM a master conv weight (sub-pixel phase weights of the up blocks): called
        after each optimizer step (the fused Adam kernel rewrites the plain packings itself)"""
        for l in self.conv_layers():
            for d in l.derived:
                d.derive()

    def attach_grads(self):
        """points every parameter's .grad at its slice of the arena (conv weights: strided OIHW views)"""
        for blk in self.modules():
            if isinstance(blk, Conv2dBlock) and blk._layer is not None:
                gw, gb = blk._layer.grad_views()
                blk.conv.weight.grad = gw
                blk.conv.bias.grad = gb
        for p, (off, numel) in getattr(self, "_dense_grads", []):
            p.grad = self._arena.view(off, numel).view(p.shape)

    def refresh_grads(self):
        """conv layers whose packed gradient layout has a negative kw stride (final conv) expose `.grad` as a copy:
        re-materialise it after a backward pass (the fused Adam kernel reads the arena directly, not this copy)"""
        for blk in self.modules():
            if isinstance(blk, Conv2dBlock) and blk._layer is not None and blk._layer.aff[blk._layer.layout][4] < 0:
                blk.conv.weight.grad = blk._layer.grad_views()[0]

    def _reserve_dense(self, p):
        if not hasattr(self, "_dense_grads"):
            self._dense_grads = []
        off = self._arena.reserve(p.numel())
        self._dense_grads.append((p, (off, p.numel())))
        return off

    def _grad_of(self, p):
        for q, (off, numel) in self._dense_grads:
            if q is p:
                return self._arena.view(off, numel).view(p.shape)
        raise KeyError


# ======================================================================================================
# Generator
# ======================================================================================================
class AdaINGen(_EngineNet):
    """AdaIN auto-encoder (reference networks.py:112-171)"""

    def __init__(self, input_dim, params):
        super().__init__()
        dim, style_dim = params["dim"], params["style_dim"]
        n_down, n_res = params["n_downsample"], params["n_res"]
        activ, pad_type = params["activ"], params["pad_type"]
        self.enc_style = StyleEncoder(4, input_dim, dim, style_dim, norm="none", activ=activ, pad_type=pad_type)
        self.enc_content = ContentEncoder(n_down, n_res, input_dim, dim, "in", activ, pad_type=pad_type)
        self.dec = Decoder(n_down, n_res, self.enc_content.output_dim, params["output_dim"], res_norm="adain",
                           activ=activ, pad_type=pad_type)
        self.mlp = MLP(style_dim, self.get_num_adain_params(self.dec), params["mlp_dim"], 3, norm="none", activ=activ)
        self.activ = activ

    # ---- reference-compatible helpers ---------------------------------------------------------------
    def get_num_adain_params(self, model):
        return sum(2 * m.num_features for m in model.modules() if isinstance(m, AdaptiveInstanceNorm2d))

    def assign_adain_params(self, adain_params, model):
        """networks.py:154-163: per AdaIN module (modules() order) first C columns -> bias, next C -> weight"""
        for m in model.modules():
            if isinstance(m, AdaptiveInstanceNorm2d):
                c = m.num_features
                m.bias = adain_params[:, :c].contiguous().view(-1)
                m.weight = adain_params[:, c:2 * c].contiguous().view(-1)
                if adain_params.size(1) > 2 * c:
                    adain_params = adain_params[:, 2 * c:]

    # ---- engine binding -----------------------------------------------------------------------------
    def _build_layers(self):
        eng, ar = self._eng, self._arena
        # registration order == optimizer order is irrelevant for the arena; windows mark the small-C layers
        for i, blk in enumerate(self.enc_style.model):
            if isinstance(blk, Conv2dBlock):
                blk.layer(eng, ar, N.WINDOW_IN if i == 0 else N.WINDOW_NONE)
        for i, blk in enumerate(self.enc_content.model):
            if isinstance(blk, Conv2dBlock):
                blk.layer(eng, ar, N.WINDOW_IN if i == 0 else N.WINDOW_NONE)
        for rb in self.enc_content.model[-1].model:
            for blk in rb.model:
                blk.layer(eng, ar)
        dec = list(self.dec.model)
        for rb in dec[0].model:
            for blk in rb.model:
                blk.layer(eng, ar)
        self.subpixel = os.environ.get("ACLGAN_SUBPIXEL", "1") != "0"
        for blk in dec[1:-1]:
            if isinstance(blk, Conv2dBlock):
                lay = blk.layer(eng, ar)
                # nearest 2x upsample + 5x5 conv in sub-pixel form (9 instead of 25 taps per output pixel, the up-sampled plane
                # is never materialised): derived phase weights around the 5x5 layer (engine.UpConvLayer)
                blk._up = E.UpConvLayer(eng, lay) if (self.subpixel and lay.k == 5 and lay.cout % 8 == 0) else None
                self._reserve_dense(blk.norm.gamma)
                self._reserve_dense(blk.norm.beta)
        dec[-1].layer(eng, ar, N.WINDOW_OUT if dec[-1].spec["cout"] <= 8 else N.WINDOW_NONE)
        head = self.enc_style.model[-1]
        for p in (head.weight, head.bias):
            self._reserve_dense(p)
        for lb in self.mlp.model:
            self._reserve_dense(lb.fc.weight)
            self._reserve_dense(lb.fc.bias)

    # ---- engine API ---------------------------------------------------------------------------------
    def enc_content_fwd(self, tape, img):
        """ContentEncoder (networks.py:230-245) on an ImgT -> content plane (pad 1: it feeds 3x3 convs)"""
        self._ensure_bound()
        eng, tw = self._eng, self.train_weights and tape.enabled
        act = _ACT[self.activ]
        m = list(self.enc_content.model)
        x = eng.pack_image(tape, img, 3, 8)
        convs = m[:-1]
        for i, blk in enumerate(convs):
            nxt_pad = convs[i + 1].spec["pad"] if i + 1 < len(convs) else 1
            x = eng.conv_block(tape, blk._layer, x, norm=N.NORM_IN, act=act, out_pad=nxt_pad, train_w=tw)
        for rb in m[-1].model:
            h = eng.conv_block(tape, rb.model[0]._layer, x, norm=N.NORM_IN, act=act, out_pad=1, train_w=tw)
            x = eng.conv_block(tape, rb.model[1]._layer, h, norm=N.NORM_IN, act=N.ACT_NONE, out_pad=1, res=x,
                               train_w=tw)
        return x

    def enc_style_fwd(self, tape, img):
        """StyleEncoder (networks.py:212-228) -> E.ImgT holding the [N, style_dim] code"""
        self._ensure_bound()
        eng, tw = self._eng, self.train_weights and tape.enabled
        act = _ACT[self.activ]
        m = list(self.enc_style.model)
        convs = [b for b in m if isinstance(b, Conv2dBlock)]
        x = eng.pack_image(tape, img, 3, 8)
        for i, blk in enumerate(convs):
            nxt_pad = convs[i + 1].spec["pad"] if i + 1 < len(convs) else 0
            x = eng.conv_block(tape, blk._layer, x, norm=N.NORM_NONE, act=act, out_pad=nxt_pad, train_w=tw)
        head = m[-1]
        c, sd = x.c_valid, head.weight.shape[0]
        L = N.lib()
        a = N.StyleHeadArgs()
        a.x, a.c_valid, a.style_dim = x.struct(), c, sd
        a.weight, a.bias = head.weight.data_ptr(), head.bias.data_ptr()
        pooled = torch.empty((x.n, c), dtype=torch.float32, device=eng.device)
        st = torch.empty((x.n, sd), dtype=torch.float32, device=eng.device)
        a.pooled, a.style = pooled.data_ptr(), st.data_ptr()
        N.check(L.aclgan_style_head_fwd(C.byref(a), E._sp()), "style_head_fwd")      # AdaptiveAvgPool2d(1) + Conv2d 1x1
        style = E.ImgT(st, requires_grad=tape.enabled)
        if tape.enabled:
            def bwd():
                if style.grad is None:
                    return
                ds = style.grad.contiguous()
                style.grad = None
                g = torch.empty((x.n, x.h, x.w, x.c), dtype=eng.prec.dtype, device=eng.device)
                a.pooled = pooled.data_ptr()        # (the closure keeps the forward's pooled activations alive)
                a.dstyle, a.gr, a.g_kind = ds.data_ptr(), g.data_ptr(), eng.prec.kind
                if tw:
                    a.dweight, a.dbias = self._grad_of(head.weight).data_ptr(), self._grad_of(head.bias).data_ptr()
                N.check(L.aclgan_style_head_bwd(C.byref(a), E._sp()), "style_head_bwd")
                x.add_gr(g)
            tape.push(bwd)
        return style

    def mlp_fwd(self, tape, style):
        """MLP (networks.py:280-292) on an E.ImgT [N, style_dim] -> E.ImgT [N, n_adain]"""
        self._ensure_bound()
        tw = self.train_weights and tape.enabled
        eng = self._eng
        fcs = [lb.fc for lb in self.mlp.model]
        x0 = style.t.reshape(style.t.shape[0], -1).contiguous()
        n = x0.shape[0]
        a = N.MlpArgs()
        a.n, a.n_layers = n, len(fcs)
        hs = [x0]
        for i, fc in enumerate(fcs):
            a.dims[i], a.dims[i + 1] = fc.weight.shape[1], fc.weight.shape[0]
            a.w[i], a.b[i] = fc.weight.data_ptr(), fc.bias.data_ptr()
            hs.append(torch.empty((n, fc.weight.shape[0]), dtype=torch.float32, device=eng.device))
        for i, h in enumerate(hs):
            a.h[i] = h.data_ptr()
        a.h0_stride = x0.shape[1]
        N.check(N.lib().aclgan_mlp_fwd(C.byref(a), E._sp()), "mlp_fwd")
        out = E.ImgT(hs[-1], requires_grad=tape.enabled)
        if tape.enabled:
            def bwd():
                if out.grad is None:
                    return
                g = out.grad.contiguous()
                out.grad = None
                scratch = [torch.empty_like(h) for h in hs[1:-1]]
                a.dh[len(fcs)] = g.data_ptr()
                for i, t in enumerate(scratch):
                    a.dh[i + 1] = t.data_ptr()
                d0 = None
                if style.requires_grad:
                    d0 = torch.empty_like(x0)
                    a.dh[0] = d0.data_ptr()
                if tw:
                    for i, fc in enumerate(fcs):
                        a.dw[i], a.db[i] = self._grad_of(fc.weight).data_ptr(), self._grad_of(fc.bias).data_ptr()
                N.check(N.lib().aclgan_mlp_bwd(C.byref(a), E._sp()), "mlp_bwd")
                if d0 is not None:
                    style.add_grad(d0.view_as(style.t))
            tape.push(bwd)
        return out

    def dec_fwd(self, tape, content, style):
        """AdaINGen.decode + Decoder.forward (networks.py:147-152, 247-264) -> E.ImgT image"""
        self._ensure_bound()
        eng, tw = self._eng, self.train_weights and tape.enabled
        act = _ACT[self.activ]
        ap = self.mlp_fwd(tape, style)
        m = list(self.dec.model)
        c = content.c_valid
        n = content.n
        d_ap = None
        if tape.enabled:
            d_ap = E.zeros(ap.t.shape, ap.t.dtype, ap.t.device)

            def bwd_ap():
                ap.add_grad(d_ap)
            tape.push(bwd_ap)
        x = content
        i_adain = 0
        for rb in m[0].model:
            for j, blk in enumerate(rb.model):
                # networks.py:154-163: first c columns of the module's slice -> bias, next c -> weight; read (and their
                # gradients written) in place, as strided views of the MLP output row / of its gradient
                lo = i_adain * 2 * c
                bias, weight = ap.t[:, lo:lo + c], ap.t[:, lo + c:lo + 2 * c]
                d_bias = d_ap[:, lo:lo + c] if d_ap is not None else None
                d_weight = d_ap[:, lo + c:lo + 2 * c] if d_ap is not None else None
                i_adain += 1
                last_rb = rb is m[0].model[-1] and j == 1
                if j == 0:
                    h = eng.conv_block(tape, blk._layer, x, norm=N.NORM_ADAIN, act=act, out_pad=1,
                                       adain=(weight, bias, d_weight, d_bias), train_w=tw)
                else:
                    up = 2 if (last_rb and len(m) > 2) else 1
                    nxt = m[2].spec["pad"] if (last_rb and len(m) > 2) else (m[-1].spec["pad"] if last_rb else 1)
                    if up == 2 and getattr(m[2], "_up", None) is not None:
                        up, nxt = 1, 1      # sub-pixel up block: it reads the SOURCE plane (reflect pad 1)
                    x = eng.conv_block(tape, blk._layer, h, norm=N.NORM_ADAIN, act=N.ACT_NONE, out_pad=nxt,
                                       upsample=up, res=x, adain=(weight, bias, d_weight, d_bias), train_w=tw)
        ups = [b for b in m[1:-1] if isinstance(b, Conv2dBlock)]
        for i, blk in enumerate(ups):
            last = i + 1 == len(ups)
            nxt = m[-1].spec["pad"] if last else ups[i + 1].spec["pad"]
            ln = (blk.norm.gamma.detach(), blk.norm.beta.detach(),
                  self._grad_of(blk.norm.gamma), self._grad_of(blk.norm.beta))
            nxt_up = None if last else getattr(ups[i + 1], "_up", None)
            if getattr(blk, "_up", None) is not None:
                x = eng.conv_block_up(tape, blk._up, x, act=act, out_pad=(1 if nxt_up is not None else nxt), ln=ln, train_w=tw)
                if not last and nxt_up is None:
                    raise N.NativeError("mixed sub-pixel / materialised up blocks are not supported")
            else:
                x = eng.conv_block(tape, blk._layer, x, norm=N.NORM_LN, act=act, out_pad=nxt,
                                   upsample=1 if last else 2, ln=ln, train_w=tw)
        return eng.conv_to_image(tape, m[-1]._layer, x, act=N.ACT_TANH, train_w=tw)

    # ---- reference-compatible tensor API (forward values only) ---------------------------------------
    def _content_from_tensor(self, content):
        """NCHW fp32 content code -> reflect-padded plane feeding the decoder's first 3x3 conv (pack kernel)"""
        eng = self._eng
        n, c, h, w = content.shape
        a = E.ActT(eng, n, h, w, c, 1)
        src = content.detach().float().contiguous()
        st = a.struct()
        N.check(N.lib().aclgan_pack_nchw(src.data_ptr(), c, C.byref(st), E._sp()), "pack_nchw")
        return a

    def _encode_impl(self, images):
        tape = E.Tape(enabled=False)
        img = E.ImgT(images)
        style = self.enc_style_fwd(tape, img)
        content = self.enc_content_fwd(tape, img)
        out = torch.empty((content.n, content.c_valid, content.h, content.w), dtype=torch.float32, device=images.device)
        st = content.struct()
        N.check(N.lib().aclgan_unpack_plane(C.byref(st), content.c_valid, out.data_ptr(), E._sp()), "unpack_plane")
        return out, style.t

    def _decode_impl(self, content, style):
        tape = E.Tape(enabled=False)
        return (self.dec_fwd(tape, self._content_from_tensor(content), E.ImgT(style)).t,)

    def encode(self, images):
        """networks.py:141-145 -> (content [N,C,H/4,W/4], style [N,style_dim,1,1])"""
        x = images.detach().float().contiguous()
        content, style = self._graphed(("encode", tuple(x.shape)), self._encode_impl, [x])
        return content, style.view(style.shape[0], -1, 1, 1)

    def decode(self, content, style):
        """networks.py:147-152 -> images [N,output_dim,H,W]"""
        c = content.detach().float().contiguous()
        st = style.detach().float().reshape(style.shape[0], -1).contiguous()
        return self._graphed(("decode", tuple(c.shape)), self._decode_impl, [c, st])[0]

    def forward(self, images):
        content, style = self.encode(images)
        return self.decode(content, style)


class VAEGen(nn.Module):
    """dead code in the reference (imported by trainer.py:4, never instantiated - SURVEY 2.1)"""

    def __init__(self, input_dim, params):
        super().__init__()
        raise NotImplementedError("VAEGen is not part of the ACL-GAN hot path")


class Vgg16(nn.Module):
    """option surface only: vgg_w is 0 in every shipped config (trainer.py:55)"""

    def __init__(self):
        super().__init__()
        raise NotImplementedError("the VGG perceptual loss is outside the B200 hot path")


# ======================================================================================================
# Discriminator
# ======================================================================================================
class MsImageDis(_EngineNet):
    """multi-scale PatchGAN (reference networks.py:21-106)"""

    def __init__(self, input_dim, params):
        super().__init__()
        self.n_layer, self.gan_type, self.dim = params["n_layer"], params["gan_type"], params["dim"]
        self.norm, self.activ = params["norm"], params["activ"]
        self.num_scales, self.pad_type = params["num_scales"], params["pad_type"]
        self.input_dim = input_dim
        self.downsample = nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)
        self.cnns = nn.ModuleList([self._make_net() for _ in range(self.num_scales)])
        if self.norm != "none":
            raise NotImplementedError("discriminator norm %r is outside the B200 hot path" % self.norm)

    def _make_net(self):
        dim = self.dim
        layers = [Conv2dBlock(self.input_dim, dim, 4, 2, 1, norm="none", activation=self.activ, pad_type=self.pad_type)]
        for _ in range(self.n_layer - 1):
            layers.append(Conv2dBlock(dim, dim * 2, 4, 2, 1, norm=self.norm, activation=self.activ,
                                      pad_type=self.pad_type))
            dim *= 2
        layers.append(nn.Conv2d(dim, 1, 1, 1, 0))
        return nn.Sequential(*layers)

    def _build_layers(self):
        for net in self.cnns:
            for i, blk in enumerate(net):
                if isinstance(blk, Conv2dBlock):
                    blk.layer(self._eng, self._arena, N.WINDOW_IN if i == 0 else N.WINDOW_NONE)
                else:
                    self._reserve_dense(blk.weight)
                    self._reserve_dense(blk.bias)

    # ---- engine API ---------------------------------------------------------------------------------
    def dis(self, tape, img0, img1=None, lsgan=None):
        """forward over all scales; img1 = second image of a channel-concatenated pair (dis_2).
        returns a list of E.ImgT logits [N,1,h,w]"""
        pyr = self.dis_pyramid(tape, img0, img1)
        return [self.dis_scale(tape, s, pyr[s], lsgan) for s in range(len(self.cnns))]

    def dis_pyramid(self, tape, img0, img1=None):
        """the input of every scale: the image(s) and their AvgPool2d(3, 2, 1) pyramid (networks.py:33,53)"""
        self._ensure_bound()
        imgs = [img0, img1]
        pyr = [imgs]
        for _ in range(len(self.cnns) - 1):
            imgs = [self._pool(tape, im) if im is not None else None for im in imgs]
            pyr.append(imgs)
        return pyr

    def dis_scale(self, tape, s, imgs, lsgan=None):
        """one scale's PatchGAN (networks.py:38-47): the scales are independent chains once the pyramid exists.
        lsgan = dict(acc=<double accumulators>, targets, weights, slots): the LSGAN terms of networks.py:67,83,98 and their
        gradient seed are fused into the head kernel (image groups along the batch, one entry per group)"""
        self._ensure_bound()
        eng, tw = self._eng, self.train_weights and tape.enabled
        act = _ACT[self.activ]
        net = self.cnns[s]
        x = eng.pack_image(tape, imgs[0], 1, 16, imgs[1])
        blocks = [b for b in net if isinstance(b, Conv2dBlock)]
        for i, blk in enumerate(blocks):
            x = eng.conv_block(tape, blk._layer, x, norm=N.NORM_NONE, act=act,
                               out_pad=1 if i + 1 < len(blocks) else 0, train_w=tw)
        return self._head(tape, net[-1], x, tw, lsgan)

    def _pool(self, tape, img):
        """AvgPool2d(3, 2, padding 1, count_include_pad=False) on the NCHW image (networks.py:33,53)"""
        n, c, h, w = img.t.shape
        ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        src = img.t.contiguous()
        t = torch.empty((n, c, ho, wo), dtype=torch.float32, device=src.device)
        a = N.AvgPoolArgs(src.data_ptr(), t.data_ptr(), n * c, h, w, 0)
        N.check(N.lib().aclgan_avgpool3x3s2_fwd(C.byref(a), E._sp()), "avgpool_fwd")
        out = E.ImgT(t, requires_grad=img.requires_grad)
        if tape.enabled and img.requires_grad:
            def bwd():
                if out.grad is None:
                    return
                go = out.grad.contiguous()
                out.grad = None
                g, acc = img.grad_buffer()
                b = N.AvgPoolArgs(go.data_ptr(), g.data_ptr(), n * c, h, w, acc)
                N.check(N.lib().aclgan_avgpool3x3s2_bwd(C.byref(b), E._sp()), "avgpool_bwd")
            tape.push(bwd)
        return out

    def _head(self, tape, conv, x, tw, lsgan=None):
        """1x1 conv dim*8 -> 1 (networks.py:45) as a warp-per-pixel dot product on the un-padded plane, fused with the GAN terms
        (per image group; LSGAN or, for gan_type 'nsgan', sigmoid + binary cross entropy) and d loss / d logits when `lsgan`
        (targets / weights / accumulator slots) is given"""
        eng = self._eng
        L = N.lib()
        c = x.c_valid
        a = N.DisHeadArgs()
        a.x, a.c_valid = x.struct(), c
        a.weight, a.bias = conv.weight.data_ptr(), conv.bias.data_ptr()
        lt = torch.empty((x.n, 1, x.h, x.w), dtype=torch.float32, device=eng.device)
        a.logits = lt.data_ptr()
        dl = None
        a.groups = 1
        if lsgan is not None:
            k = len(lsgan["targets"])
            a.groups = k
            for i in range(k):
                a.target[i], a.gweight[i], a.loss_slot[i] = lsgan["targets"][i], lsgan["weights"][i], lsgan["slots"][i]
            a.loss = lsgan["acc"].data_ptr()
            self._check_gan()            # (like the reference, an unknown gan_type only fails where a loss is formed)
            a.gan_kind = _GAN_KIND[self.gan_type]
            if tape.enabled:
                dl = torch.empty((x.n, 1, x.h, x.w), dtype=torch.float32, device=eng.device)
                a.dlogits = dl.data_ptr()
        N.check(L.aclgan_dis_head_fwd(C.byref(a), E._sp()), "dis_head_fwd")
        logit = E.ImgT(lt, requires_grad=tape.enabled)
        if tape.enabled:
            def bwd():
                d = dl
                if logit.grad is not None:       # a gradient seeded by the caller (tests / custom losses) adds to the fused seed
                    # (seeded on the caller's stream, consumed on this chain's stream: keep the allocator from reusing it early)
                    if logit.grad.is_cuda:
                        logit.grad.record_stream(torch.cuda.current_stream())
                    d = logit.grad.contiguous() if d is None else E.acc_(d, logit.grad.contiguous())
                    logit.grad = None
                if d is None:
                    return
                b = N.DisHeadBwdArgs()
                b.x, b.c_valid, b.g_kind = x.struct(), c, eng.prec.kind
                b.weight, b.dlogits = conv.weight.data_ptr(), d.data_ptr()
                if tw:
                    b.dweight, b.dbias = self._grad_of(conv.weight).data_ptr(), self._grad_of(conv.bias).data_ptr()
                g = None
                if x.requires_grad:
                    g = torch.empty((x.n, x.h, x.w, x.c), dtype=eng.prec.dtype, device=eng.device)
                    b.gr = g.data_ptr()
                N.check(L.aclgan_dis_head_bwd(C.byref(b), E._sp()), "dis_head_bwd")
                if g is not None:
                    x.add_gr(g)
            tape.push(bwd)
        return logit

    # ---- reference-compatible tensor API (forward values only) ---------------------------------------
    def forward(self, x):
        tape = E.Tape(enabled=False)
        return [o.t for o in self.dis(tape, E.ImgT(x.detach().float()))]

    def _lsgan(self, outs, target):
        """the per-scale GAN terms of networks.py:60-106 on logit tensors (the updates use the fused head kernel instead)"""
        if self.gan_type == "nsgan":
            return sum(F.binary_cross_entropy(torch.sigmoid(o), torch.full_like(o, target)) for o in outs)
        return sum(torch.mean((o - target) ** 2) for o in outs)

    def _check_gan(self):
        if self.gan_type not in _GAN_KIND:
            assert 0, "Unsupported GAN type: {}".format(self.gan_type)

    def calc_dis_loss(self, input_fake, input_real):
        self._check_gan()
        return self._lsgan(self.forward(input_fake), 0.0) + self._lsgan(self.forward(input_real), 1.0)

    def calc_gen_loss(self, input_fake):
        self._check_gan()
        return self._lsgan(self.forward(input_fake), 1.0)

    def calc_gen_d2_loss(self, input_fake, input_real):
        self._check_gan()
        return self._lsgan(self.forward(input_fake), 1.0) + self._lsgan(self.forward(input_real), 0.0)
