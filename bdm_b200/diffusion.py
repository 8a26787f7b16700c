"""Diffusion schedules and the three BDM sampling procedures around the denoisers.

  DDPMSchedule   the PC^2 side: diffusers 0.21.0 `DDPMScheduler(beta_start=1e-5, beta_end=8e-3,
                 beta_schedule='linear', clip_sample=False)` (reference model/model.py:51-62,
                 config/structured.py:105-107).  diffusers is not vendored: restated from its published
                 algorithm (epsilon prediction, `fixed_small` variance clamped at 1e-20).  PARITY UNPINNED.
  PVDSchedule    the prior side: the in-repo GaussianDiffusion (reference pvd/__init__.py:18-68
                 coefficients, :136-224 p_mean_variance / p_sample with model_mean_type='eps',
                 model_var_type='fixedsmall', betas linear 1e-4..0.02 `:477`).
  BDMSampler     vanilla PC^2 sampling (model.py:123-214), BDM-Blending (main_blending.py:186-347) and
                 BDM-Merging (main_merging.py:369-523) with the shipped schedule
                 roll_step=16, milestones=[1000,968,936,872,128,64,32,0] as default.

Host-side schedule math only; every per-step tensor op runs on the device.
"""
import numpy as np
import torch


class DDPMSchedule:
    def __init__(self, beta_start=1e-5, beta_end=8e-3, num_train_timesteps=1000):
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.timesteps = list(range(num_train_timesteps - 1, -1, -1))  # set_timesteps(1000)

    def coefficients(self, t):
        """Python floats for one step t -> (x0_from_xt, x0_from_eps, coef_x0, coef_xt, sigma)"""
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[t - 1] if t > 0 else torch.tensor(1.0)
        beta_prod_t, beta_prod_prev = 1 - a_t, 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        coef_x0 = (a_prev ** 0.5 * cur_beta) / beta_prod_t
        coef_xt = cur_alpha ** 0.5 * beta_prod_prev / beta_prod_t
        var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_beta, min=1e-20)
        sigma = var ** 0.5 if t > 0 else torch.tensor(0.0)
        return float(a_t ** 0.5), float(beta_prod_t ** 0.5), float(coef_x0), float(coef_xt), float(sigma)

    def step(self, eps, t, x_t, generator=None):
        """x_{t-1} from the predicted noise (scheduler.step(...).prev_sample)"""
        sqrt_a, sqrt_b, coef_x0, coef_xt, sigma = self.coefficients(int(t))
        x0 = (x_t - sqrt_b * eps) / sqrt_a
        prev = coef_x0 * x0 + coef_xt * x_t
        if int(t) > 0:
            prev = prev + sigma * torch.randn(eps.shape, generator=generator, device=eps.device, dtype=eps.dtype)
        return prev


class PVDSchedule:
    def __init__(self, b_start=1e-4, b_end=0.02, time_num=1000):
        betas = np.linspace(b_start, b_end, time_num).astype(np.float64)
        alphas = 1.0 - betas
        ac = torch.from_numpy(np.cumprod(alphas, axis=0)).float()
        ac_prev = torch.from_numpy(np.append(1.0, ac[:-1].numpy())).float()
        b32, a32 = torch.from_numpy(betas).float(), torch.from_numpy(alphas).float()
        self.num_timesteps = time_num
        self.sqrt_recip_ac = torch.sqrt(1.0 / ac)
        self.sqrt_recipm1_ac = torch.sqrt(1.0 / ac - 1)
        post_var = b32 * (1.0 - ac_prev) / (1.0 - ac)
        self.post_log_var = torch.log(torch.max(post_var, 1e-20 * torch.ones_like(post_var)))
        self.coef1 = b32 * torch.sqrt(ac_prev) / (1.0 - ac)
        self.coef2 = (1.0 - ac_prev) * torch.sqrt(a32) / (1.0 - ac)

    def step(self, eps, t, x_t, generator=None):
        t = int(t)
        x0 = float(self.sqrt_recip_ac[t]) * x_t - float(self.sqrt_recipm1_ac[t]) * eps
        mean = float(self.coef1[t]) * x0 + float(self.coef2[t]) * x_t
        noise = torch.randn(x_t.shape, generator=generator, device=x_t.device, dtype=x_t.dtype)
        if t == 0:
            return mean
        return mean + float(torch.exp(0.5 * self.post_log_var[t])) * noise


class GraphedStep:
    """CUDA-graph capture of `eps = denoiser(conditioning(x_t), t)` for one batch shape.

    A denoiser step is several hundred small launches (92 native-op calls of ours plus every dense
    layer); at 16 shapes per GPU the host cannot keep the device fed.  The whole noise prediction is
    captured once into a CUDA graph with static input/output buffers and replayed per step; the DDPM
    update (which draws fresh noise from the caller's generator) stays outside.  Results are the same
    kernels on the same data, so bit-identical to the eager path."""

    def __init__(self, fn, x_example, t_example, warmup=3):
        import torch
        self.x = x_example.clone()
        self.t = t_example.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                fn(self.x, self.t)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = fn(self.x, self.t)

    def __call__(self, x, t):
        self.x.copy_(x)
        self.t.copy_(t)
        self.graph.replay()
        return self.out


DEFAULT_MILESTONES = (1000, 968, 936, 872, 128, 64, 32, 0)
DEFAULT_ROLL_STEP = 16


class BDMSampler:
    """Holds the co-resident networks of one rank and runs the coupled chains for a batch of shapes.

    pc2_net:  PointCloudModel  ((B,N,3+C), t) -> (B,N,3)
    pvd_net:  PVCNN2_PVD       ((B,3,N), t)   -> (B,3,N)        (optional)
    fuse_net: PVCNNFuse                                          (optional, Merging only)
    conditioner: projection.ProjectionConditioner for the batch
    """

    def __init__(self, pc2_net, conditioner, pvd_net=None, fuse_net=None, generator=None):
        self.pc2_net, self.pvd_net, self.fuse_net = pc2_net, pvd_net, fuse_net
        self.cond = conditioner
        self.gen = generator
        self.ddpm = DDPMSchedule()
        self.pvd = PVDSchedule()
        self.forwards = dict(pc2=0, pvd=0, fuse=0)
        self._graphs = {}

    def enable_cuda_graphs(self, x_example):
        """Capture the PC^2 (and PVD, if present) noise predictions for clouds shaped like `x_example`
        (B,N,3).  Later steps with that shape replay the graph; other shapes run eagerly."""
        import torch
        b = x_example.shape[0]
        tt = torch.full((b,), 500, device=x_example.device, dtype=torch.long)
        self._graphs[("pc2", tuple(x_example.shape))] = GraphedStep(self._pc2_eps, x_example, tt)
        if self.pvd_net is not None:
            x_cf = x_example.permute(0, 2, 1).contiguous()
            self._graphs[("pvd", tuple(x_cf.shape))] = GraphedStep(lambda x, t: self.pvd_net(x, t), x_cf, tt)

    def _pc2_eps(self, x_t, tt):
        """noise prediction of the PC^2 branch: conditioning + denoiser.  With a channel-last feature map
        and a CUDA conditioner the projected features are written directly in the denoiser's
        channel-first layout (one pass instead of gather + concat + transpose)."""
        fused = getattr(self.cond, "channel_last", False) and hasattr(self.cond, "get_input_channel_first") \
            and hasattr(self.pc2_net, "forward_channel_first") and x_t.is_cuda
        if fused:
            return self.pc2_net.forward_channel_first(self.cond.get_input_channel_first(x_t), tt)
        return self.pc2_net(self.cond.get_input_with_conditioning(x_t), tt)

    # -- one denoising step of each kind ---------------------------------------------------------
    def pc2_step(self, x_t, t):
        b = x_t.shape[0]
        tt = torch.full((b,), int(t), device=x_t.device, dtype=torch.long)
        graphed = self._graphs.get(("pc2", tuple(x_t.shape)))
        eps = graphed(x_t, tt) if graphed is not None else self._pc2_eps(x_t, tt)
        self.forwards['pc2'] += 1
        return self.ddpm.step(eps, t, x_t, self.gen)

    def pvd_step(self, x_t_cf, t):
        """x_t_cf channel-first (B,3,N)"""
        b = x_t_cf.shape[0]
        tt = torch.full((b,), int(t), device=x_t_cf.device, dtype=torch.long)
        graphed = self._graphs.get(("pvd", tuple(x_t_cf.shape)))
        eps = graphed(x_t_cf, tt) if graphed is not None else self.pvd_net(x_t_cf, tt)
        self.forwards['pvd'] += 1
        return self.pvd.step(eps, t, x_t_cf, self.gen)

    # -- chains ----------------------------------------------------------------------------------
    def pc2_chain(self, x, start_time, end_time):
        """model.py:216-289 interaction_sample: timesteps[1000-start : 1000-end] = start-1 ... end"""
        for t in range(start_time - 1, end_time - 1, -1):
            x = self.pc2_step(x, t)
        return x

    def pvd_chain(self, x, start_time, final_time):
        """pvd/__init__.py:450-473 generate_pvd_xyz on (B,N,3) clouds (main_blending.py:176-183)"""
        x = x.permute(0, 2, 1).float().contiguous()
        for t in reversed(range(final_time, start_time)):
            x = self.pvd_step(x, t)
        return x.permute(0, 2, 1)

    def fuse_step(self, from_prior, from_recon, timestep):
        """model.py:510-570 nstep_fuse"""
        from_prior = from_prior - from_prior.mean(dim=1, keepdim=True)
        from_recon = from_recon - from_recon.mean(dim=1, keepdim=True)
        b = from_recon.shape[0]
        tt = torch.full((b,), int(timestep), device=from_recon.device, dtype=torch.long)
        cond_in = self.cond.get_input_with_conditioning(from_recon)
        eps = self.fuse_net(cond_in.transpose(1, 2), from_prior.transpose(1, 2).contiguous(), tt).transpose(1, 2)
        self.forwards['fuse'] += 1
        return self.ddpm.step(eps, timestep, from_recon, self.gen)

    def _init_cloud(self, b, n, device, centre):
        x = torch.randn(b, n, 3, generator=self.gen, device=device)
        return x - x.mean(dim=1, keepdim=True) if centre else x

    # -- the three procedures ----------------------------------------------------------------------
    @torch.no_grad()
    def sample_vanilla(self, b, n, device, num_steps=1000):
        return self.pc2_chain(self._init_cloud(b, n, device, centre=False), num_steps, 0)

    @torch.no_grad()
    def sample_blending(self, b, n, device, milestones=DEFAULT_MILESTONES, roll_step=DEFAULT_ROLL_STEP,
                        mask_generator=None):
        m = list(milestones)
        x = self._init_cloud(b, n, device, centre=True)
        for i in range(len(m) - 1):
            if i == 0:
                x = self.pc2_chain(x, m[0], m[1] - roll_step)
            elif i == len(m) - 2:
                x = self.pc2_chain(x, m[i] - roll_step, m[i + 1])
            else:
                x = self.pc2_chain(x, m[i] - roll_step, m[i + 1])
                recon = self.pc2_chain(x.clone(), m[i + 1], m[i + 1] - roll_step)
                prior = self.pvd_chain(x.clone(), m[i + 1], m[i + 1] - roll_step)
                # per-point coin flip on the host generator (main_blending.py:330-344)
                pick = torch.randint(0, 2, (b, n), generator=mask_generator).to(device).bool()
                x = torch.where(pick.unsqueeze(-1), prior, recon)
        return x

    @torch.no_grad()
    def sample_merging(self, b, n, device, milestones=DEFAULT_MILESTONES, roll_step=DEFAULT_ROLL_STEP):
        m = list(milestones)
        x = self._init_cloud(b, n, device, centre=True)
        for i in range(len(m) - 1):
            if i == 0:
                x = self.pc2_chain(x, m[0], m[1] - roll_step)
            elif i == len(m) - 2:
                x = self.pc2_chain(x, m[i] - roll_step, m[i + 1])
            else:
                x = self.pc2_chain(x, m[i] - roll_step, m[i + 1])
                recon = self.pc2_chain(x.clone(), m[i + 1], m[i + 1] - roll_step + 1)
                prior = self.pvd_chain(x.clone(), m[i + 1], m[i + 1] - roll_step + 1)
                x = self.fuse_step(prior, recon, m[i + 1] - roll_step)
        return x


def forward_counts(milestones=DEFAULT_MILESTONES, roll_step=DEFAULT_ROLL_STEP, mode="merging"):
    """Denoiser forwards per shape implied by a schedule: used by the benchmark to
    turn a measured step time into shapes/s without running all 1000 steps."""
    m = list(milestones)
    pc2 = pvd = fuse = 0
    for i in range(len(m) - 1):
        if i == 0:
            pc2 += m[0] - (m[1] - roll_step)
        elif i == len(m) - 2:
            pc2 += (m[i] - roll_step) - m[i + 1]
        else:
            pc2 += (m[i] - roll_step) - m[i + 1]
            branch = roll_step if mode == "blending" else roll_step - 1
            pc2 += branch
            pvd += branch
            fuse += 0 if mode == "blending" else 1
    return dict(pc2=pc2, pvd=pvd, fuse=fuse)
