"""Host-side mirror of the reference's model wrapper for the rescaling path: `create_model(opt)` /
`SelfCModel` (models/__init__.py:5-15, models/SelfC_model.py:27-319, models/base_model.py:77-107), inference half.

Same call sequence as test_rescaling.py uses -- feed_data(data) -> test() -> get_current_visuals() -- and the same
checkpoint handling (state_dict file, optional 'module.' prefix, strict load).  Differences, on purpose:
  * test() runs the network ONCE per GOP; the reference's extra padded pass whose result is discarded
    (SelfC_model.py:203-243) is not reproduced;
  * no DataParallel wrapper: one process per GPU (the wrapper's `.module` attribute is provided for compatibility);
  * training: optimize_parameters (SelfC_model.py:148-183) runs forward + backward + clip + Adam through the C-ABI
    (selfc_b200/train.py), fp32 or bf16x3 mode (`precision`); the gradient all-reduce replaces DistributedDataParallel.
"""
from __future__ import annotations

import logging
from collections import OrderedDict

import torch

from . import engine as _engine
from . import networks
from .global_var import GlobalVar
from .sharding import GOP

logger = logging.getLogger("base")


class _ModuleHandle:
    """`model.netG.module` compatibility with code written against the DataParallel wrapper."""

    def __init__(self, net):
        self.module = net

    def __getattr__(self, k):
        return getattr(self.module, k)

    def __call__(self, *a, **k):
        return self.module(*a, **k)


class SelfCModel:
    def __init__(self, opt):
        self.opt = opt
        if not torch.cuda.is_available():
            raise RuntimeError("selfc_b200.model.SelfCModel needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.is_train = bool(opt["is_train"])
        net = networks.define_G(opt).to(self.device)
        self.netG = _ModuleHandle(net)
        self.load()
        self.gop = GOP
        self.log_dict = OrderedDict()
        if self.is_train:
            net.train()
            to = opt["train"]
            from .train import Trainer
            self.train_opt = to
            self.trainer = Trainer(net, self.device, lr=float(to["lr_G"]), betas=(float(to["beta1"]), float(to["beta2"])),
                                   weight_decay=float(to["weight_decay_G"] or 0.0),
                                   max_norm=float(to["gradient_clipping"]) if to["gradient_clipping"] else None)
            self.cur_lr = float(to["lr_G"])
        else:
            net.eval()

    # ---- checkpoint handling: base_model.py:77-107 ----------------------------------------------------------
    def load(self):
        path = (self.opt["path"] or {}).get("pretrain_model_G") if self.opt["path"] else None
        if path:
            logger.info("Loading model for G [%s] ...", path)
            self.load_network(path, self.netG.module, bool(self.opt["path"].get("strict_load", True)))

    @staticmethod
    def load_network(load_path, network, strict=True):
        load_net = torch.load(load_path, map_location="cpu")
        clean = OrderedDict()
        for k, v in load_net.items():
            if "Quantization_H265_Suggrogate" in k:
                continue
            clean[k[7:] if k.startswith("module.") else k] = v
        network.load_state_dict(clean, strict=strict)

    def save_network(self, save_path):
        sd = OrderedDict((k, v.detach().cpu()) for k, v in self.netG.module.state_dict().items())
        torch.save(sd, save_path)

    def save(self, iter_label):
        """SelfC_model.py:318-319 / base_model.py:77-85: `<models>/<iter>_G.pth`."""
        import os
        os.makedirs(self.opt["path"]["models"], exist_ok=True)
        path = os.path.join(self.opt["path"]["models"], "{}_{}.pth".format(iter_label, "G"))
        self.save_network(path)
        return path

    # ---- training state: base_model.py:109-133 -----------------------------------------------------------------
    def training_state(self, epoch, iter_step):
        """The dict the reference saves: epoch, iter, one scheduler and one optimizer state.  The optimizer entry has
        torch.optim.Adam's state_dict layout (per-parameter step / exp_avg / exp_avg_sq in named_parameters order), so the
        file loads into the reference's own optimizer as well."""
        tr = self.trainer
        offs = tr._offsets.tolist()
        state = {}
        for i, p in enumerate(tr.params):
            state[i] = {"step": torch.tensor(float(tr.step_count)),
                        "exp_avg": tr.m[offs[i]:offs[i + 1]].view(p.shape).detach().cpu().clone(),
                        "exp_avg_sq": tr.v[offs[i]:offs[i + 1]].view(p.shape).detach().cpu().clone()}
        group = {"lr": self.cur_lr, "betas": tuple(tr.betas), "eps": tr.eps, "weight_decay": tr.weight_decay, "amsgrad": False,
                 "initial_lr": float(self.train_opt["lr_G"]), "params": list(range(len(tr.params)))}
        sched = {"last_epoch": int(iter_step), "milestones": list(self.train_opt["lr_steps"] or []),
                 "gamma": float(self.train_opt["lr_gamma"] or 1.0)}
        net = self.netG.module
        return {"epoch": epoch, "iter": iter_step, "schedulers": [sched], "optimizers": [{"state": state, "param_groups": [group]}],
                # the counter-based noise stream of the sampler (not in the reference, whose noise is the unseeded CUDA generator)
                "noise": {"seed": int(net.noise_seed), "offset": int(net.noise_offset)}}

    def save_training_state(self, epoch, iter_step):
        import os
        os.makedirs(self.opt["path"]["training_state"], exist_ok=True)
        path = os.path.join(self.opt["path"]["training_state"], "{}.state".format(iter_step))
        torch.save(self.training_state(epoch, iter_step), path)
        return path

    def resume_training(self, resume_state):
        """Restore the Adam moments, step count and learning rate.  (The reference's body is commented out,
        base_model.py:122-133, so its resume silently restarts the optimiser; the network weights come from
        path.pretrain_model_G in both.)"""
        tr = self.trainer
        opt_state = resume_state["optimizers"][0]
        offs = tr._offsets.tolist()
        if len(opt_state["state"]) != len(tr.params):
            raise ValueError("Wrong lengths of optimizers")
        steps = set()
        for i, p in enumerate(tr.params):
            st = opt_state["state"][i]
            tr.m[offs[i]:offs[i + 1]].view(p.shape).copy_(st["exp_avg"])
            tr.v[offs[i]:offs[i + 1]].view(p.shape).copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) != 1:
            raise ValueError("per-parameter Adam step counts differ")
        tr.step_count = steps.pop()
        self.cur_lr = float(opt_state["param_groups"][0]["lr"])
        noise = resume_state.get("noise")
        if noise is not None:                  # continue the eps stream where the saved run stopped
            self.netG.module.set_noise(int(noise["seed"]), int(noise["offset"]))

    # ---- feed_data: SelfC_model.py:93-132 -----------------------------------------------------------------
    def feed_data(self, data):
        real = data["GT"]                                   # [B,3,t,H,W]
        t_len = GlobalVar.get_Temporal_LEN()
        clip_length = real.size(2)
        if clip_length < t_len:                             # pad short clips with copies of the last frame
            pads = real[:, :, -1:].expand(-1, -1, t_len - clip_length, -1, -1)
            real = torch.cat([real, pads], dim=2)
        real = real.to(self.device, non_blocking=True)
        self.real_H = real.transpose(1, 2).reshape(-1, 3, real.size(3), real.size(4)).contiguous()
        dist = self.opt["distortion"]
        if "LQ" in data:
            lq = data["LQ"].to(self.device)
            self.ref_L = lq.transpose(1, 2).reshape(-1, 3, lq.size(3), lq.size(4))
        elif dist == "sr_bd":
            self.ref_L = _engine.gaussian_downsample(self.real_H)
        elif dist == "pytorch_bicubic":
            self.ref_L = torch.nn.functional.interpolate(self.real_H, scale_factor=1.0 / self.opt["scale"], mode="area")
        else:
            raise NotImplementedError(f"distortion {dist!r} is not on the rescaling path (only sr_bd / pytorch_bicubic)")
        return clip_length

    # ---- optimize_parameters: SelfC_model.py:148-183 ---------------------------------------------------------
    def optimize_parameters(self, step):
        """zero_grad, forward (down, quantise, up), the two losses x 144*144*3, backward, gradient clipping, Adam."""
        if not self.is_train:
            raise RuntimeError("optimize_parameters needs is_train (a training options file)")
        to = self.train_opt
        for key, want in (("pixel_criterion_forw", "l2"), ("pixel_criterion_back", "l1")):
            if (to[key] or want) != want:
                raise NotImplementedError(f"{key}={to[key]!r}: the CUDA training step implements the SelfC-large YAML's losses (l2 / l1)")
        if float(to["lambda_fit_forw"] or 1) != 1.0 or float(to["lambda_rec_back"] or 1) != 1.0:
            raise NotImplementedError("lambda_fit_forw / lambda_rec_back other than 1 are not implemented")
        t = GlobalVar.get_Temporal_LEN()
        net = self.netG.module
        losses = self.trainer.step(self.real_H, self.ref_L.contiguous(), t, eps=net._eps_override, seed=net.noise_seed,
                                   offset=net.noise_offset, lr=self.cur_lr)
        net.noise_offset += 1
        total, l_forw, l_back = [float(v) for v in losses.cpu()]
        self.log_dict["l_forw_fit"] = l_forw
        self.log_dict["l_back_rec"] = l_back
        self.log_dict["loss_c"] = 0.0
        self.log_dict["loss"] = total

    def update_learning_rate(self, cur_iter, warmup_iter=-1):
        """base_model.py:40-58 with the MultiStepLR scheme of the training YAML."""
        from .train import multistep_lr
        to = self.train_opt
        lr = multistep_lr(float(to["lr_G"]), int(cur_iter), to["lr_steps"] or [], float(to["lr_gamma"] or 1.0))
        if warmup_iter and warmup_iter > 0 and cur_iter < warmup_iter:
            lr = lr / warmup_iter * cur_iter
        self.cur_lr = lr

    def get_current_learning_rate(self):
        return self.cur_lr

    def get_current_log(self):
        return self.log_dict

    # ---- test: SelfC_model.py:185-250 -----------------------------------------------------------------------
    def test(self):
        net = self.netG.module
        t = GlobalVar.get_Temporal_LEN()
        bt, c, hh, ww = self.real_H.shape
        b = bt // t
        clips = self.real_H.reshape(b, t, c, hh, ww)
        saved_t = t
        forw_L, forw_H, fake_H, sample_H = [], [], [], []
        with torch.no_grad():
            for g0 in range(0, t, self.gop):
                ids = list(range(g0, min(t, g0 + self.gop)))
                real = len(ids)
                ids += [t - 1] * (self.gop - real)
                x = clips[:, ids].reshape(b * self.gop, c, hh, ww)
                GlobalVar.set_Temporal_LEN(self.gop)
                out, _ = net(x=x)
                lr = net_quantize(out[:, :3])
                hr, hf = net(x=lr, rev=True)
                for lst, ten, cc in ((forw_L, lr, 3), (forw_H, out[:, 3:], 48), (fake_H, hr[:, :3], 3), (sample_H, hf, 48)):
                    v = ten.reshape(b, self.gop, cc, ten.shape[-2], ten.shape[-1])[:, :real]
                    lst.append(v)
        GlobalVar.set_Temporal_LEN(saved_t)
        cat = lambda lst: torch.cat(lst, dim=1).reshape(-1, *lst[0].shape[2:])
        self.forw_L, self.forw_H, self.fake_H, self.sample_H = cat(forw_L), cat(forw_H), cat(fake_H), cat(sample_H)

    def get_current_visuals(self):
        out = OrderedDict()
        out["SR"] = self.fake_H.detach()
        out["LR_ref"] = self.ref_L.detach()
        out["LR"] = self.forw_L.detach()
        out["GT"] = self.real_H.detach()
        out["forw_H"] = self.forw_H.detach()
        return out


def net_quantize(x: torch.Tensor) -> torch.Tensor:
    """Quantization() of the reference (Quantization.py:19-26) on the CUDA path."""
    return _engine.quantize(x.contiguous())[1]


def create_model(opt):
    """models/__init__.py:5-15, rescaling branch."""
    model = opt["model"]
    if model in ("SelfC_GMM",):
        m = SelfCModel(opt)
    else:
        raise NotImplementedError(f"Model [{model}] not recognized by selfc_b200 (only SelfC_GMM).")
    logger.info("Model [%s] is created.", m.__class__.__name__)
    return m
