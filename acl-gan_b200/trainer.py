"""Drop-in `trainer` module: `aclgan_Trainer` with the reference's surface (trainer.py:14-332 of the reference:
same child modules / attribute names / loss_* attributes / checkpoint files), running the adversarial-
consistency cycle on the CUDA engine with a hand-scheduled backward pass instead of autograd.

Differences that do not change any result (SURVEY.md 8a "work the reference performs that does not affect
any result"): dis_update runs the generators without recording a backward pass (the reference back-propagates
into them and zeroes those grads at trainer.py:91), gen_update computes no discriminator weight gradients
(zeroed at :248), style encoders whose output is unused are skipped, dis_A(x_a) is evaluated once with
weight 1 instead of twice with weight 1/2.
"""
import os

import torch
import torch.nn as nn

import aclgan_native as N
import engine as E
from networks import AdaINGen, MsImageDis, get_engine
from utils import get_model_list, get_scheduler, weights_init


class aclgan_Trainer(nn.Module):
    def __init__(self, hyperparameters):
        super().__init__()
        hp = hyperparameters
        lr = hp["lr"]
        # construction order == reference (trainer.py:19-25): it fixes the RNG stream and the state_dict
        self.gen_AB = AdaINGen(hp["input_dim_a"], hp["gen"])
        self.gen_BA = AdaINGen(hp["input_dim_a"], hp["gen"])
        self.dis_A = MsImageDis(hp["input_dim_a"], hp["dis"])
        self.dis_B = MsImageDis(hp["input_dim_a"], hp["dis"])
        self.dis_2 = MsImageDis(hp["input_dim_b"], hp["dis"])
        self.instancenorm = nn.InstanceNorm2d(512, affine=False)
        self.style_dim = hp["gen"]["style_dim"]
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        display_size = int(hp["display_size"])
        self.z_1 = torch.randn(display_size, self.style_dim, 1, 1).to(dev)
        self.z_2 = torch.randn(display_size, self.style_dim, 1, 1).to(dev)
        self.z_3 = torch.randn(display_size, self.style_dim, 1, 1).to(dev)

        beta1, beta2 = hp["beta1"], hp["beta2"]
        dis_params = list(self.dis_A.parameters()) + list(self.dis_B.parameters()) + list(self.dis_2.parameters())
        gen_params = list(self.gen_AB.parameters()) + list(self.gen_BA.parameters())
        self.dis_opt = torch.optim.Adam([p for p in dis_params if p.requires_grad], lr=lr, betas=(beta1, beta2),
                                        weight_decay=hp["weight_decay"])
        self.gen_opt = torch.optim.Adam([p for p in gen_params if p.requires_grad], lr=lr, betas=(beta1, beta2),
                                        weight_decay=hp["weight_decay"])
        self.dis_scheduler = get_scheduler(self.dis_opt, hp)
        self.gen_scheduler = get_scheduler(self.gen_opt, hp)
        self.alpha = hp["alpha"]
        self.focus_lam = hp["focus_loss"]

        self.apply(weights_init(hp["init"]))
        self.dis_A.apply(weights_init("gaussian"))
        self.dis_B.apply(weights_init("gaussian"))
        self.dis_2.apply(weights_init("gaussian"))

        if hp.get("vgg_w", 0) > 0:
            raise NotImplementedError("vgg_w > 0: the VGG perceptual loss is outside the B200 hot path")
        self.precision = hp.get("precision", os.environ.get("ACLGAN_PRECISION", "bf16"))
        self._hp = hp
        for net in (self.gen_AB, self.gen_BA, self.dis_A, self.dis_B, self.dis_2):
            net.precision_hint = self.precision         # stand-alone gen.encode / decode (test.py) run in the trainer's mode
        self.use_graphs = bool(int(hp.get("cuda_graphs", os.environ.get("ACLGAN_CUDA_GRAPHS", "1"))))
        self._graphs = {}
        self._launches = {}
        self._lplans = {}
        self.merge_passes = bool(int(hp.get("merge_passes", os.environ.get("ACLGAN_MERGE_PASSES", "1"))))
        self.parallel_dis = bool(int(hp.get("parallel_dis", os.environ.get("ACLGAN_PARALLEL_DIS", "1"))))
        # one chain per (discriminator, scale) instead of per discriminator: the sub-wave kernels of the small scales
        # overlap nine-fold.  (It exposed a cross-stream allocator hazard - a gradient seeded on the caller's stream and
        # released inside a chain's closure was reused early; fixed with record_stream at the two hand-off points and
        # guarded by tests/test_gpu_step.py::test_schedule_variants_agree.)
        self.parallel_scales = bool(int(hp.get("parallel_scales", os.environ.get("ACLGAN_PARALLEL_SCALES", "1"))))
        self._side_streams = None
        self._early_stream = None
        # dis_update (graph replay + all-reduce + Adam) runs on its own stream; the next gen_update starts its generator
        # passes right away and waits for it (an external event-wait node inside its graph) only before the discriminator
        # passes.  Results of dis_update (loss_dis_*, discriminator gradients / weights) are ordered after
        # torch.cuda.synchronize(), the next gen_update, or any read of a loss_dis_* attribute.
        self.overlap_updates = bool(int(hp.get("overlap_updates", os.environ.get("ACLGAN_OVERLAP_UPDATES", "1"))))
        self._dis_stream = None
        self._dis_event = None
        self.expose_grads = bool(int(hp.get("expose_grads", 1)))   # keep every param.grad readable after an update
        self._ready = False
        self._noise = None          # optional injected style noise (tests / graph replay): list of 3 tensors

    # ------------------------------------------------------------------------------------------ dis_update ordering
    def _join_dis(self):
        """orders the caller's stream after the last dis_update (no-op when nothing is pending)"""
        if self._dis_event is not None and torch.cuda.is_available():
            torch.cuda.current_stream().wait_event(self._dis_event)

    def _wait_dis_in_graph(self):
        if self._dis_event is not None:
            N.check(N.lib().aclgan_stream_wait_external_event(E._sp(), self._dis_event.cuda_event), "plan:wait_external_event")

    # ------------------------------------------------------------------------------------------ engine
    def _setup(self):
        if self._ready:
            return
        eng = get_engine(self.precision)
        self.eng = eng
        self.gen_arena = E.GradArena(eng.device)
        self.dis_arena = E.GradArena(eng.device)
        for net in (self.gen_AB, self.gen_BA):
            net.bind(eng, self.gen_arena)
        for net in (self.dis_A, self.dis_B, self.dis_2):
            net.bind(eng, self.dis_arena)
        for net in self._nets():
            net._ensure_bound()
        self.gen_arena.finalize()
        self.dis_arena.finalize()
        for net in self._nets():
            net.attach_grads()
        self._adam_gen = self._adam_group(self.gen_opt, (self.gen_AB, self.gen_BA), self.gen_arena)
        self._adam_dis = self._adam_group(self.dis_opt, (self.dis_A, self.dis_B, self.dis_2), self.dis_arena)
        if self.overlap_updates and self.use_graphs and torch.cuda.is_available():
            self._dis_stream = torch.cuda.Stream()
            self._dis_event = torch.cuda.Event()
            self._dis_event.record()            # creates the event handle the gen graph's wait node refers to
        self._ready = True

    def _nets(self):
        return (self.gen_AB, self.gen_BA, self.dis_A, self.dis_B, self.dis_2)

    def _repack_dirty(self):
        """re-derives the packed bf16 weights of every layer whose fp32 master was replaced outside the fused Adam kernel
        (load_state_dict hook, or a manual edit followed by net.mark_dirty()); cheap no-op otherwise"""
        for net in self._nets():
            for l in net.conv_layers():
                if l.dirty:
                    l.repack()

    # ------------------------------------------------------------------------------------------ optimizer
    def _adam_group(self, opt, nets, arena):
        """device tables for the fused Adam kernel; the torch.optim.Adam object keeps owning the state tensors
        (exp_avg / exp_avg_sq / step) so optimizer.pt stays interchangeable with the reference's"""
        import ctypes as C
        import networks as NW
        dev = self.eng.device
        layer_of = {}
        for net in nets:
            for blk in net.modules():
                if isinstance(blk, NW.Conv2dBlock) and blk._layer is not None:
                    layer_of[id(blk.conv.weight)] = blk._layer
        params = opt.param_groups[0]["params"]
        table = (N.AdamTensor * len(params))()
        chunks = []
        step0 = 0.0
        for i, p in enumerate(params):
            st = opt.state[p]
            if "exp_avg" not in st:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            for k in ("exp_avg", "exp_avg_sq"):
                if st[k].device != p.device or not st[k].is_contiguous():
                    st[k] = st[k].to(p.device).contiguous()
            step0 = max(step0, float(st["step"]))
            t = table[i]
            t.p, t.m, t.v = p.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            t.g = arena.flat.data_ptr()
            shape = list(p.shape)
            while len(shape) < 4:
                shape = [1] + shape
            lay = layer_of.get(id(p))
            if lay is not None:
                off, strides = lay.grad_layout()
                t.goff = off
                for j in range(4):
                    t.gs[j] = strides[j]
                for k in (0, 1):
                    for pl in range(self.eng.prec.planes):
                        t.pk[k][pl] = lay.packed[k][pl].data_ptr()
                    for j in range(5):
                        t.aff[k][j] = lay.aff[k][j]
            else:
                g = p.grad
                assert g is not None and g.is_contiguous() and g.untyped_storage().data_ptr() == arena.flat.untyped_storage().data_ptr()
                t.goff = g.storage_offset()
                acc = 1
                for j in reversed(range(4)):
                    t.gs[j] = acc
                    acc *= shape[j]
            for j in range(4):
                t.d[j] = shape[j]
            t.planes = self.eng.prec.planes
            for c in range(N.lib().aclgan_adam_units(C.byref(t))):
                chunks += [i, c]
        hp = self._hp
        world = 1
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size()
        hyper = torch.tensor([opt.param_groups[0]["lr"], hp["beta1"], hp["beta2"], 1e-8, hp["weight_decay"],
                              1.0 / world, step0, 0.0], dtype=torch.float32, device=dev)
        raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8).to(dev)
        return dict(table=raw, chunks=torch.tensor(chunks, dtype=torch.int32, device=dev), hyper=hyper,
                    n_chunks=len(chunks) // 2, opt=opt, params=params, step=step0)

    def _adam_step(self, grp):
        L = N.lib()
        sp = E._sp()
        N.check(L.aclgan_adam_advance(grp["hyper"].data_ptr(), sp), "adam_advance")
        N.check(L.aclgan_adam_step(grp["table"].data_ptr(), grp["chunks"].data_ptr(), grp["n_chunks"],
                                   grp["hyper"].data_ptr(), sp), "adam_step")
        grp["step"] += 1
        grp["opt"]._opt_called = True

    def _sync_opt_state(self):
        """mirror the device-side step counters into the torch optimizer state (before state_dict())"""
        if not self._ready:
            return
        self._join_dis()
        for grp in (self._adam_gen, self._adam_dis):
            n = float(grp["hyper"][6].item())
            for p in grp["params"]:
                grp["opt"].state[p]["step"] = torch.tensor(n)

    def _draw_noise(self, n):
        """three CPU randn draws moved to the device, exactly as trainer.py:99-101 / 254-256"""
        if self._noise is not None:
            zs, self._noise = self._noise, None
            return [E.ImgT(z.to(self.eng.device).float().reshape(n, self.style_dim)) for z in zs]
        return [E.ImgT(torch.randn(n, self.style_dim, 1, 1).to(self.eng.device).reshape(n, self.style_dim))
                for _ in range(3)]

    def _allreduce(self, arena):
        """data parallel: sum the flat gradient buffer over ranks (NCCL); the 1/world scale is folded into Adam"""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(arena.flat)

    # ------------------------------------------------------------------------------------------ pieces
    def recon_criterion(self, input, target):
        return torch.mean(torch.abs(input - target))

    def focus_translation(self, x_fg, x_bg, x_focus):
        x_map = ((x_focus + 1) / 2).repeat(1, 3, 1, 1)
        return x_fg * x_map + x_bg * (1 - x_map)

    def _blend(self, tape, out, bg):
        """decoder output [N,4,H,W] -> focus-blended RGB (trainer.py:85-88,108-111) by the blend kernel; the raw mask stays
        channel 3 of `out`"""
        import ctypes as C
        n, _, h, w = out.t.shape
        o4, bgt = out.t.contiguous(), bg.t.contiguous()
        dst = torch.empty((n, 3, h, w), dtype=torch.float32, device=o4.device)
        a = N.BlendArgs()
        a.out4, a.bg, a.dst, a.n, a.h, a.w = o4.data_ptr(), bgt.data_ptr(), dst.data_ptr(), n, h, w
        N.check(N.lib().aclgan_focus_blend_fwd(C.byref(a), E._sp()), "focus_blend_fwd")
        res = E.ImgT(dst, requires_grad=out.requires_grad or bg.requires_grad)
        if tape.enabled and res.requires_grad:
            def bwd():
                if res.grad is None:
                    return
                d = res.grad.contiguous()
                res.grad = None
                a.ddst = d.data_ptr()
                if out.requires_grad:
                    g, a.acc_out4 = out.grad_buffer()
                    a.dout4 = g.data_ptr()
                if bg.requires_grad:
                    g, a.acc_bg = bg.grad_buffer()
                    a.dbg = g.data_ptr()
                N.check(N.lib().aclgan_focus_blend_bwd(C.byref(a), E._sp()), "focus_blend_bwd")
            tape.push(bwd)
        return res

    # ------------------------------------------------------------------------------------------ stream forks
    def _dis_parallel(self, tape, jobs):
        """runs the three discriminator passes `jobs` = [callable(tape) -> logits] as independent chains: forward on side
        streams forked from the current one, and (through the tape tags) backward likewise.  The passes only read shared
        tensors and write private ones; gradients meet again in the image nodes after the join."""
        if not self.parallel_dis or not torch.cuda.is_available():
            return [job(tape) for job in jobs]
        if self._side_streams is None:
            self._side_streams = []
        while len(self._side_streams) < len(jobs):
            self._side_streams.append(torch.cuda.Stream())
        main = torch.cuda.current_stream()
        outs = []
        for k, job in enumerate(jobs):
            st = self._side_streams[k]
            st.wait_stream(main)
            tape.tag = k
            try:
                with torch.cuda.stream(st):
                    outs.append(job(tape))
            finally:
                tape.tag = None
        for st in self._side_streams[:len(jobs)]:
            main.wait_stream(st)
        return outs

    def _dis_all(self, tape, inputs):
        """inputs: [(discriminator, img0, img1 | None, lsgan spec | None)] -> list of per-discriminator logit lists.  The
        pooling pyramids are built first; then every (discriminator, scale) PatchGAN runs as its own parallel chain.  The
        LSGAN terms and their gradient seeds are computed by the head kernel of every scale (networks._head)."""
        if not self.parallel_scales:
            return self._dis_parallel(tape, [lambda t, d=d, a=a, b=b, ls=ls: d.dis(t, a, b, ls) for d, a, b, ls in inputs])
        pyrs = [d.dis_pyramid(tape, a, b) for d, a, b, _ in inputs]
        jobs, index = [], []
        for i, (d, _, _, ls) in enumerate(inputs):
            for sc in range(len(d.cnns)):
                jobs.append(lambda t, d=d, sc=sc, imgs=pyrs[i][sc], ls=ls: d.dis_scale(t, sc, imgs, ls))
                index.append(i)
        flat = self._dis_parallel(tape, jobs)
        outs = [[] for _ in inputs]
        for i, o in zip(index, flat):
            outs[i].append(o)
        return outs

    def _cat(self, tape, imgs):
        """batch-concatenates image nodes so a discriminator runs ONCE over all of them (3x fewer, 3x larger launches)"""
        if len(imgs) == 1:
            return imgs[0]
        res = E.ImgT(E.cat0([i.t for i in imgs]), requires_grad=any(i.requires_grad for i in imgs))
        if tape.enabled and res.requires_grad:
            def bwd():
                if res.grad is None:
                    return
                off = 0
                for i in imgs:
                    n = i.t.shape[0]
                    if i.requires_grad:
                        i.add_grad(res.grad[off:off + n])
                    off += n
                res.grad = None
            tape.push(bwd)
        return res

    def _slice(self, tape, img, lo, hi):
        """rows [lo, hi) of a batched image node (inverse of _cat)"""
        res = E.ImgT(img.t[lo:hi], requires_grad=img.requires_grad)
        if tape.enabled and res.requires_grad:
            def bwd():
                if res.grad is None:
                    return
                E.acc_(img.zero_grad_buffer()[lo:hi], res.grad.contiguous())
                res.grad = None
            tape.push(bwd)
        return res

    @staticmethod
    def _lsgan(outs, target, weight):
        """sum over scales of mean((o - t)^2) (networks.py:67,83,98); seeds d loss / d logits scaled by `weight`.
        Plain-torch helper for tests / custom losses on top of MsImageDis.dis(); the updates use the fused head kernel."""
        total = 0
        for o in outs:
            diff = o.t - target
            total = total + torch.mean(diff * diff)
            if o.requires_grad and weight != 0:
                o.add_grad(diff * (2.0 * weight / diff.numel()))
        return total

    @staticmethod
    def _scaled(z, alpha):
        """alpha * z as an image node (trainer.py:109,119,264,269: alpha multiplies only z_2)"""
        t = z.t.contiguous()
        return E.ImgT(E.axpby(torch.empty_like(t), t, None, alpha, 0.0))

    def _cycle(self, tape, x_a, x_b, zs, need_recon, early=None):
        """encode / decode cycle shared by both updates (trainer.py:103-133 and 258-280); `early(x_B_fake)` is called as
        soon as the first translation exists (dis_update starts its dis_B pass there, next to the second translation)"""
        focus = self.focus_lam > 0
        AB, BA = self.gen_AB, self.gen_BA
        z_1, z_2, z_3 = zs
        r = {}
        if need_recon and self.merge_passes:
            # passes that share weights and do not depend on each other run as ONE batched pass (every layer on the path
            # is per-sample, so the results are those of the separate passes): AB encodes [x_a; x_b] and decodes
            # [(c_1, z_1); (c_4, s_4)], BA decodes c_2 with [alpha z_2; s_2] - half the launches, better-filled grids.
            # The AB and the BA pass are independent chains (own weights, own gradients): forked like the discriminators.
            n = x_a.t.shape[0]
            xab = self._cat(tape, [x_a, x_b])
            az_2 = self._scaled(z_2, self.alpha)

            def chain_ab(t):
                c_14 = AB.enc_content_fwd(t, xab)
                s_4 = AB.enc_style_fwd(t, x_b)
                return AB.dec_fwd(t, c_14, self._cat(t, [z_1, s_4]))

            def chain_ba(t):
                c_2 = BA.enc_content_fwd(t, x_a)
                s_2 = BA.enc_style_fwd(t, x_a)
                return BA.dec_fwd(t, self.eng.dup_plane(t, c_2), self._cat(t, [az_2, s_2]))

            o_14, o_22 = self._dis_parallel(tape, [chain_ab, chain_ba])
            o_b, r["o_rec_b"] = self._slice(tape, o_14, 0, n), self._slice(tape, o_14, n, 2 * n)
            o_a, r["o_rec_a"] = self._slice(tape, o_22, 0, n), self._slice(tape, o_22, n, 2 * n)
        elif not need_recon:
            az_2 = self._scaled(z_2, self.alpha)
            o_b, o_a = self._dis_parallel(tape, [lambda t: AB.dec_fwd(t, AB.enc_content_fwd(t, x_a), z_1),
                                                 lambda t: BA.dec_fwd(t, BA.enc_content_fwd(t, x_a), az_2)])
        else:
            c_1 = AB.enc_content_fwd(tape, x_a)
            c_2 = BA.enc_content_fwd(tape, x_a)
            o_b = AB.dec_fwd(tape, c_1, z_1)
            o_a = BA.dec_fwd(tape, c_2, self._scaled(z_2, self.alpha))
            if need_recon:
                s_2 = BA.enc_style_fwd(tape, x_a)
                c_4 = AB.enc_content_fwd(tape, x_b)
                s_4 = AB.enc_style_fwd(tape, x_b)
                r["o_rec_a"] = BA.dec_fwd(tape, c_2, s_2)
                r["o_rec_b"] = AB.dec_fwd(tape, c_4, s_4)
        if focus:
            x_B_fake = self._blend(tape, o_b, x_a)
            x_A_fake = self._blend(tape, o_a, x_a)
        else:
            x_B_fake, x_A_fake = o_b, o_a
        if early is not None:
            early(x_B_fake)
        c_3 = BA.enc_content_fwd(tape, x_B_fake)
        o_a2 = BA.dec_fwd(tape, c_3, z_3)
        x_A2_fake = self._blend(tape, o_a2, x_B_fake) if focus else o_a2
        r.update(o_b=o_b, o_a=o_a, o_a2=o_a2, x_B_fake=x_B_fake, x_A_fake=x_A_fake, x_A2_fake=x_A2_fake)
        self._last_cycle = r        # handles only (tests / image dumps read them); freed by the next update
        return r

    # ------------------------------------------------------------------------------------------ updates
    def gen_update(self, x_a, x_b, hyperparameters):
        self._setup()
        if self.use_graphs:
            self._replay("gen", x_a, x_b, hyperparameters)
        else:
            self._repack_dirty()
            self._gen_fwd_bwd(x_a, x_b, hyperparameters, self._draw_noise(x_a.size(0)))
        self._allreduce(self.gen_arena)
        self._adam_step(self._adam_gen)
        self.gen_AB.derive_weights()
        self.gen_BA.derive_weights()
        if self.expose_grads:
            self.gen_AB.refresh_grads()
            self.gen_BA.refresh_grads()

    # ------------------------------------------------------------------------------------------ loss bookkeeping
    # Every loss term is reduced on the device into a double accumulator (`acc`, one memset per update); ONE kernel then forms
    # all loss_* scalars and the weighted totals as out = M @ acc (M is built on the host outside any graph capture).
    GEN_SLOTS = dict(advA1=0, advA2=1, advB=2, adv2a=3, adv2b=4, idtA=17, idtB=18)        # focus tag t: S1, S2, digit, size = 5+4t ..
    GEN_OUT = ("loss_gen_adv_A", "loss_gen_adv_B", "loss_gen_adv_2", "loss_gen_focus_B_size", "loss_gen_focus_B_digit",
               "loss_gen_focus_A_size", "loss_gen_focus_A_digit", "loss_gen_focus_A2_size", "loss_gen_focus_A2_digit",
               "loss_idt_A", "loss_idt_B", "loss_gen_total")
    DIS_OUT = ("loss_dis_A", "loss_dis_B", "loss_dis_2", "loss_dis_total")

    def _loss_plan(self, kind, hp, shape):
        """accumulators + combination matrix of one update kind; cached per (kind, weights, input shape)"""
        key = (kind, tuple(shape), tuple(sorted((k, v) for k, v in hp.items() if isinstance(v, (int, float)))))
        plan = self._lplans.get(key)
        if plan is not None:
            return plan
        dev = self.eng.device
        gw, gcw = float(hp["gan_w"]), float(hp["gan_cw"])
        if kind == "gen":
            names, K = self.GEN_OUT, 19
            M = torch.zeros((len(names), K), dtype=torch.float32)
            row = {n: i for i, n in enumerate(names)}
            S = self.GEN_SLOTS
            M[row["loss_gen_adv_A"], S["advA1"]] = M[row["loss_gen_adv_A"], S["advA2"]] = 0.5
            M[row["loss_gen_adv_B"], S["advB"]] = 1.0
            M[row["loss_gen_adv_2"], S["adv2a"]] = M[row["loss_gen_adv_2"], S["adv2b"]] = 1.0
            M[row["loss_idt_A"], S["idtA"]] = M[row["loss_idt_B"], S["idtB"]] = 1.0
            for t, tag in enumerate(("B", "A", "A2")):
                M[row["loss_gen_focus_%s_digit" % tag], 5 + 4 * t + 2] = 1.0
                M[row["loss_gen_focus_%s_size" % tag], 5 + 4 * t + 3] = 1.0
            tot = M[row["loss_gen_total"]]
            tot += gw * M[row["loss_gen_adv_A"]] + gw * M[row["loss_gen_adv_B"]] + gcw * M[row["loss_gen_adv_2"]]
            tot += float(hp["recon_x_w"]) * (M[row["loss_idt_A"]] + M[row["loss_idt_B"]])
            if hp["focus_loss"] > 0:
                fscale = float(hp["focus_loss"]) / float(shape[2]) / float(shape[3]) / float(shape[0]) / 3.0      # trainer.py:161
                for tag in ("B", "A", "A2"):
                    tot += fscale * (M[row["loss_gen_focus_%s_size" % tag]] + M[row["loss_gen_focus_%s_digit" % tag]])
        else:
            names, K = self.DIS_OUT, 8
            M = torch.zeros((len(names), K), dtype=torch.float32)
            # slots: dis_A over [x_a, x_A_fake, x_A2_fake] -> 0, 1, 2; dis_B over [x_B_fake, x_b] -> 3, 4; dis_2 pairs -> 5, 6
            M[0, 0], M[0, 1], M[0, 2] = 1.0, 0.5, 0.5          # (la1 + la2 + 2 la0) / 2: dis_A(x_a) counted twice at 1/2
            M[1, 3] = M[1, 4] = 1.0
            M[2, 5] = M[2, 6] = 1.0
            M[3] = gw * M[0] + gw * M[1] + gcw * M[2]
        plan = dict(names=names, K=K, M=M.to(dev), acc=torch.zeros(K, dtype=torch.float64, device=dev),
                    out=torch.zeros(len(names), dtype=torch.float32, device=dev))
        self._lplans[key] = plan
        return plan

    def _loss_finish(self, plan):
        N.check(N.lib().aclgan_loss_combine(plan["acc"].data_ptr(), plan["M"].data_ptr(), plan["out"].data_ptr(),
                                            len(plan["names"]), plan["K"], E._sp()), "loss_combine")
        for i, name in enumerate(plan["names"]):
            if "focus" in name and not self.focus_lam > 0:
                continue            # the reference only creates the focus loss attributes when the branch is on (trainer.py:146)
            setattr(self, name, plan["out"][i])

    def _gen_fwd_bwd(self, x_a, x_b, hyperparameters, zs):
        import ctypes as C
        hp = hyperparameters
        L = N.lib()
        plan = self._loss_plan("gen", hp, x_a.shape)
        acc = plan["acc"]
        E.zero_(acc)
        self.gen_arena.zero_()
        if self.eng.pool is not None:
            self.eng.pool.begin()
        for d in (self.dis_A, self.dis_B, self.dis_2):
            d.train_weights = False                       # their weight grads are discarded (trainer.py:248)
        tape = E.Tape()
        n, _, h, w = x_a.shape
        xa, xb = E.ImgT(x_a.detach().float()), E.ImgT(x_b.detach().float())
        focus = hp["focus_loss"] > 0
        r = self._cycle(tape, xa, xb, zs, need_recon=True)

        gw, gcw = hp["gan_w"], hp["gan_cw"]
        S = self.GEN_SLOTS
        self._wait_dis_in_graph()           # the discriminators' weights come from the (possibly still running) dis_update
        cat_a = self._cat(tape, [r["x_A_fake"], r["x_A2_fake"]])
        cat_2a, cat_2b = self._cat(tape, [xa, xa]), self._cat(tape, [r["x_A_fake"], r["x_A2_fake"]])
        # LSGAN terms (networks.py:77-106) and d loss / d logits inside the head kernels: calc_gen_loss targets 1; calc_gen_d2_loss
        # targets (1, 0); loss_gen_adv_A is the 1/2-weighted sum of two calls (trainer.py:136-137)
        self._dis_all(tape, [
            (self.dis_A, cat_a, None, dict(acc=acc, targets=[1.0, 1.0], weights=[0.5 * gw, 0.5 * gw], slots=[S["advA1"], S["advA2"]])),
            (self.dis_B, r["x_B_fake"], None, dict(acc=acc, targets=[1.0], weights=[gw], slots=[S["advB"]])),
            (self.dis_2, cat_2a, cat_2b, dict(acc=acc, targets=[1.0, 0.0], weights=[gcw, gcw], slots=[S["adv2a"], S["adv2b"]]))])

        if focus:
            # trainer.py:146-161 (sums over the WHOLE batch; normalised by H*W*B*3): pass 1 reduces sum(m - upper),
            # sum(lower - m), sum 1/(|m - .5| + eps); pass 2 forms the size loss and writes d(size + digit)/d mask into channel 3
            gscale = 0.5 * float(hp["focus_loss"]) / float(h * w * n * 3)
            for t, out in enumerate((r["o_b"], r["o_a"], r["o_a2"])):
                o4 = out.t
                a = N.LossReduceArgs()
                a.mode, a.n, a.ca, a.c, a.h, a.w = N.LOSS_FOCUS, n, 4, 1, h, w
                a.a, a.acc, a.slot = o4.data_ptr(), acc.data_ptr(), 5 + 4 * t
                a.upper, a.lower, a.eps = hp["focus_upper"], hp["focus_lower"], hp["focus_epsilon"]
                N.check(L.aclgan_loss_reduce(C.byref(a), E._sp()), "loss_reduce(focus)")
                g = N.FocusGradArgs()
                g.out4, g.dout4 = o4.data_ptr(), out.zero_grad_buffer().data_ptr()
                g.n, g.h, g.w, g.slot, g.size_slot, g.acc = n, h, w, 5 + 4 * t, 5 + 4 * t + 3, 1
                g.sums, g.delta, g.eps, g.gscale = acc.data_ptr(), hp["focus_delta"], hp["focus_epsilon"], gscale
                N.check(L.aclgan_focus_grad(C.byref(g), E._sp()), "focus_grad")

        # identity L1 (trainer.py:61-62,162-165) on the first three channels of the reconstructions
        rw = float(hp["recon_x_w"])
        for rec, ref, slot in ((r["o_rec_a"], xa, S["idtA"]), (r["o_rec_b"], xb, S["idtB"])):
            ca = rec.t.shape[1]
            a = N.LossReduceArgs()
            a.mode, a.n, a.ca, a.c, a.h, a.w = N.LOSS_L1, n, ca, 3, h, w
            a.a, a.b, a.acc, a.slot = rec.t.data_ptr(), ref.t.data_ptr(), acc.data_ptr(), slot
            a.da, a.acc_da, a.gscale = rec.zero_grad_buffer().data_ptr(), 1, rw / float(n * 3 * h * w)
            N.check(L.aclgan_loss_reduce(C.byref(a), E._sp()), "loss_reduce(l1)")
        self._loss_finish(plan)

        tape.backward(self._side_streams)
        for d in (self.dis_A, self.dis_B, self.dis_2):
            d.train_weights = True

    def dis_update(self, x_a, x_b, hyperparameters):
        self._setup()
        if self._dis_stream is not None and self.use_graphs:
            st = self._dis_stream
            st.wait_stream(torch.cuda.current_stream())         # inputs and the generators' weights (last gen_update)
            with torch.cuda.stream(st):
                self._replay("dis", x_a, x_b, hyperparameters)
                self._allreduce(self.dis_arena)
                self._adam_step(self._adam_dis)
                self._dis_event.record(st)
            for t in (x_a, x_b):
                if t.is_cuda:
                    t.record_stream(st)
            return
        if self.use_graphs:
            self._replay("dis", x_a, x_b, hyperparameters)
        else:
            self._repack_dirty()
            self._dis_fwd_bwd(x_a, x_b, hyperparameters, self._draw_noise(x_a.size(0)))
        self._allreduce(self.dis_arena)
        self._adam_step(self._adam_dis)

    def _dis_fwd_bwd(self, x_a, x_b, hyperparameters, zs):
        hp = hyperparameters
        plan = self._loss_plan("dis", hp, x_a.shape)
        acc = plan["acc"]
        E.zero_(acc)
        self.dis_arena.zero_()
        if self.eng.pool is not None:
            self.eng.pool.begin()
        xa, xb = E.ImgT(x_a.detach().float()), E.ImgT(x_b.detach().float())
        gw, gcw = hp["gan_w"], hp["gan_cw"]
        early_stream = None

        def dis_b_pass(fake_b):
            # calc_dis_loss(fake, real) = LSGAN(fake, 0) + LSGAN(real, 1)  (networks.py:60-75)
            tb = E.Tape()
            self._dis_all(tb, [(self.dis_B, self._cat(tb, [fake_b, xb]), None,
                                dict(acc=acc, targets=[0.0, 1.0], weights=[gw, gw], slots=[3, 4]))])
            tb.backward(self._side_streams)

        def early(x_B_fake):
            # dis_B only needs x_B_fake: its whole forward + backward runs on a side stream while the generators still
            # compute the second translation (x_A2_fake) on the caller's stream
            nonlocal early_stream
            if self._early_stream is None:
                self._early_stream = torch.cuda.Stream()
            early_stream = self._early_stream
            early_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(early_stream):
                dis_b_pass(E.ImgT(x_B_fake.t))

        use_early = self.parallel_dis and torch.cuda.is_available()
        # the generators only produce the fakes here: no backward pass is recorded for them (trainer.py:91)
        r = self._cycle(E.Tape(enabled=False), xa, xb, zs, need_recon=False, early=early if use_early else None)
        fake_a, fake_a2, fake_b = (E.ImgT(r[k].t) for k in ("x_A_fake", "x_A2_fake", "x_B_fake"))
        if early_stream is not None:
            torch.cuda.current_stream().wait_stream(early_stream)
        else:
            dis_b_pass(fake_b)

        tape = E.Tape()
        # each discriminator runs once over the batch-concatenation of its inputs; dis_A(x_a) is counted twice with
        # weight 1/2 in the reference (trainer.py:283-284) == once with weight 1
        cat_a = self._cat(tape, [xa, fake_a, fake_a2])
        cat_2a, cat_2b = self._cat(tape, [xa, xa]), self._cat(tape, [fake_a, fake_a2])
        self._dis_all(tape, [
            (self.dis_A, cat_a, None, dict(acc=acc, targets=[1.0, 0.0, 0.0], weights=[gw, 0.5 * gw, 0.5 * gw], slots=[0, 1, 2])),
            (self.dis_2, cat_2a, cat_2b, dict(acc=acc, targets=[0.0, 1.0], weights=[gcw, gcw], slots=[5, 6]))])
        self._loss_finish(plan)
        tape.backward(self._side_streams)

    @property
    def launches_per_step_pair(self):
        """kernels of libaclgan_b200.so launched per dis_update + gen_update (counted while capturing the graphs)"""
        return sum(self._launches.values()) if len(self._launches) == 2 else None

    # ------------------------------------------------------------------------------------------ CUDA graphs
    def _replay(self, kind, x_a, x_b, hp):
        """forward + backward of one update as a captured CUDA graph (the ~2500 kernel launches of an update would
        otherwise be bound by host-side launch overhead).  Inputs / style noise are copied into static buffers; the
        loss_* attributes are static 0-dim tensors the replay overwrites.  The gradient all-reduce and the Adam
        kernel stay outside the graph."""
        n = x_a.size(0)
        key = (kind, tuple(x_a.shape), tuple(sorted((k, v) for k, v in hp.items() if isinstance(v, (int, float)))))
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._capture(kind, x_a, x_b, hp)
            self._graphs[key] = ent
        self._repack_dirty()
        ent["xa"].copy_(x_a, non_blocking=True)
        ent["xb"].copy_(x_b, non_blocking=True)
        # pinned noise staging is double-buffered: the H2D copy of the previous replay may still be queued behind tens of
        # milliseconds of GPU work, so a buffer is only rewritten after the copy that last read it has completed
        slot = ent["zslot"] = 1 - ent.get("zslot", 1)
        zpin = ent["zpin"][slot]
        if ent["zev"][slot] is not None:
            ent["zev"][slot].synchronize()
        if self._noise is not None:
            zs, self._noise = self._noise, None
            zpin.copy_(torch.stack([z.reshape(n, self.style_dim).float().cpu() for z in zs]))
        else:       # three CPU randn draws in the reference's order (trainer.py:99-101 / 254-256)
            for i in range(3):
                zpin[i].copy_(torch.randn(n, self.style_dim, 1, 1).view(n, self.style_dim))
        ent["z"].copy_(zpin, non_blocking=True)
        if ent["zev"][slot] is None:
            ent["zev"][slot] = torch.cuda.Event()
        ent["zev"][slot].record()
        ent["graph"].replay()
        for k, v in ent["losses"].items():
            setattr(self, k, v)
        self._last_cycle = ent["cycle"]

    def _capture(self, kind, x_a, x_b, hp):
        dev = self.eng.device
        n = x_a.size(0)
        ent = dict(xa=torch.empty(x_a.shape, dtype=torch.float32, device=dev),
                   xb=torch.empty(x_b.shape, dtype=torch.float32, device=dev),
                   z=torch.zeros((3, n, self.style_dim), dtype=torch.float32, device=dev),
                   zpin=[torch.zeros((3, n, self.style_dim), dtype=torch.float32).pin_memory() for _ in range(2)],
                   zev=[None, None])
        ent["xa"].copy_(x_a)
        ent["xb"].copy_(x_b)
        impl = self._gen_fwd_bwd if kind == "gen" else self._dis_fwd_bwd

        def run():
            self.eng.pool = ent["pool"]
            try:
                impl(ent["xa"], ent["xb"], hp, [E.ImgT(ent["z"][i]) for i in range(3)])
            finally:
                self.eng.pool = None

        # warm-up off the capture stream (lazy initialisation of kernels / cuBLAS); forward + backward only,
        # so no parameter, optimizer or RNG state is touched
        self._repack_dirty()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        ent["pool"] = E.SumsPool(dev)           # counting mode: the warm-up run measures the statistics workspace
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ent["pool"] = E.SumsPool(dev, ent["pool"].off)
        before = set(k for k in vars(self) if k.startswith("loss_"))
        graph = torch.cuda.CUDAGraph()
        n0 = N.launch_count
        with torch.cuda.graph(graph):
            run()
        self._launches[kind] = N.launch_count - n0 + 2      # + the Adam kernels launched outside the graph
        if kind == "gen":
            names = [k for k in vars(self) if k.startswith("loss_gen") or k.startswith("loss_idt")]
        else:
            names = ["loss_dis_A", "loss_dis_B", "loss_dis_2", "loss_dis_total"]
        ent["losses"] = {k: self.__dict__.get("_dv_" + k[5:], self.__dict__.get(k)) for k in names}
        ent["cycle"] = self._last_cycle
        ent["graph"] = graph
        return ent

    # ------------------------------------------------------------------------------------------ inference
    def forward(self, x_a, x_b):
        raise NotImplementedError("aclgan_Trainer.forward is dead code in the reference (crashes with the shipped "
                                  "4-channel decoder, SURVEY 2.1); use sample() / gen_*.encode / decode")

    def _sample_one(self, a, z1, z2, z3):
        """the per-image translations of trainer.py:190-226 on the engine (planes stay on the device between encode and decode)"""
        focus = self.focus_lam > 0
        tape = E.Tape(enabled=False)
        AB, BA = self.gen_AB, self.gen_BA
        xa = E.ImgT(a)
        c_1, s_1 = BA.enc_content_fwd(tape, xa), BA.enc_style_fwd(tape, xa)
        o = BA.dec_fwd(tape, c_1, E.ImgT(z1))
        rec = BA.dec_fwd(tape, c_1, s_1)
        ob = AB.dec_fwd(tape, AB.enc_content_fwd(tape, xa), E.ImgT(z2))
        xb_img = self._blend(tape, ob, xa) if focus else ob
        o2 = BA.dec_fwd(tape, BA.enc_content_fwd(tape, xb_img), E.ImgT(z3))
        if focus:
            return (self._blend(tape, o, xa).t, o.t, xb_img.t, ob.t, self._blend(tape, o2, xb_img).t, o2.t, rec.t)
        return (o.t, xb_img.t, o2.t, rec.t)

    def _recon_b(self, x_b):
        tape = E.Tape(enabled=False)
        AB = self.gen_AB
        xb = E.ImgT(x_b)
        return (AB.dec_fwd(tape, AB.enc_content_fwd(tape, xb), AB.enc_style_fwd(tape, xb)).t,)

    def sample(self, x_a, x_b):
        """trainer.py:179-245: per-image translations with the fixed display noise; every image replays ONE captured CUDA graph
        of the batch-1 forward passes (5 encodes / decodes + the focus blends)"""
        self._setup()
        self._join_dis()
        self.eval()
        focus = self.focus_lam > 0
        nets = (self.gen_AB, self.gen_BA)
        x_a, x_b = x_a.detach().float().contiguous(), x_b.detach().float().contiguous()
        sd = self.style_dim
        rows = []
        for i in range(x_a.size(0)):
            ins = [x_a[i:i + 1].contiguous()] + [z[i].reshape(1, sd).float().contiguous() for z in (self.z_1, self.z_2, self.z_3)]
            rows.append(self.gen_AB._graphed(("sample", tuple(ins[0].shape), focus), self._sample_one, ins, owners=nets))
        cols = [torch.cat(c) for c in zip(*rows)]
        self.train()
        if focus:
            fa, oa, fb, ob, fa2, oa2, rec = cols
            return (x_a, fa, oa[:, 3:4], fb, ob[:, 3:4], fa2, oa2[:, 3:4], rec[:, :3], rec[:, 3:4])
        fa, fb, fa2, rec = cols
        # the reference re-encodes the WHOLE x_b batch inside its per-image loop (trainer.py:225-226) and concatenates the
        # copies: x_B_recon has display_size x batch rows
        rb = self.gen_AB._graphed(("recon_b", tuple(x_b.shape)), self._recon_b, [x_b], owners=nets)[0]
        return (x_a, fa, fb, fa2, rec, x_b, rb.repeat(x_a.size(0), 1, 1, 1))

    # ------------------------------------------------------------------------------------------ schedule / io
    def update_learning_rate(self):
        if self.dis_scheduler is not None:
            self.dis_scheduler.step()
        if self.gen_scheduler is not None:
            self.gen_scheduler.step()
        self._join_dis()
        if self._ready:         # the fused Adam kernel reads lr from device memory (outside any captured graph)
            self._adam_gen["hyper"][0:1].fill_(self.gen_opt.param_groups[0]["lr"])
            self._adam_dis["hyper"][0:1].fill_(self.dis_opt.param_groups[0]["lr"])

    def resume(self, checkpoint_dir, hyperparameters):
        self._join_dis()
        last = get_model_list(checkpoint_dir, "gen")
        sd = torch.load(last)
        self.gen_AB.load_state_dict(sd["AB"])
        self.gen_BA.load_state_dict(sd["BA"])
        iterations = int(last[-11:-3])
        sd = torch.load(get_model_list(checkpoint_dir, "dis"))
        self.dis_A.load_state_dict(sd["A"])
        self.dis_B.load_state_dict(sd["B"])
        self.dis_2.load_state_dict(sd["2"])
        sd = torch.load(os.path.join(checkpoint_dir, "optimizer.pt"))
        self.dis_opt.load_state_dict(sd["dis"])
        self.gen_opt.load_state_dict(sd["gen"])
        self.dis_scheduler = get_scheduler(self.dis_opt, hyperparameters, iterations)
        self.gen_scheduler = get_scheduler(self.gen_opt, hyperparameters, iterations)
        if self._ready:
            # the networks keep their engine layers, gradient arenas and captured graphs (load_state_dict copied into the
            # same parameter tensors); only the optimizer state tensors were replaced: rebuild the Adam device tables
            # around them and re-derive the packed bf16 weights from the restored masters
            self._adam_gen = self._adam_group(self.gen_opt, (self.gen_AB, self.gen_BA), self.gen_arena)
            self._adam_dis = self._adam_group(self.dis_opt, (self.dis_A, self.dis_B, self.dis_2), self.dis_arena)
            self._repack_dirty()
        print("Resume from iteration %d" % iterations)
        return iterations

    def save(self, snapshot_dir, iterations):
        self._sync_opt_state()
        gen_name = os.path.join(snapshot_dir, "gen_%08d.pt" % (iterations + 1))
        dis_name = os.path.join(snapshot_dir, "dis_%08d.pt" % (iterations + 1))
        opt_name = os.path.join(snapshot_dir, "optimizer.pt")
        torch.save({"AB": self.gen_AB.state_dict(), "BA": self.gen_BA.state_dict()}, gen_name)
        torch.save({"A": self.dis_A.state_dict(), "B": self.dis_B.state_dict(), "2": self.dis_2.state_dict()}, dis_name)
        torch.save({"gen": self.gen_opt.state_dict(), "dis": self.dis_opt.state_dict()}, opt_name)


def _dis_loss_property(name):
    """loss_dis_* are produced on dis_update's own stream: reading one orders the caller's stream after that update"""
    key = "_dv_" + name[5:]          # (storage name without "loss": utils.write_loss scans attribute names for it)

    def get(self):
        if key not in self.__dict__:
            raise AttributeError(name)
        self._join_dis()
        return self.__dict__[key]

    def set_(self, value):
        self.__dict__[key] = value

    return property(get, set_)


for _n in ("loss_dis_A", "loss_dis_B", "loss_dis_2", "loss_dis_total"):
    setattr(aclgan_Trainer, _n, _dis_loss_property(_n))
