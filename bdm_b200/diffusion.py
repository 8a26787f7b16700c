"""Diffusion schedules and the three BDM sampling procedures around the denoisers.

  DDPMSchedule   the PC^2 side: diffusers 0.21.0 `DDPMScheduler(beta_start=1e-5, beta_end=8e-3,
                 beta_schedule='linear', clip_sample=False)` (reference model/model.py:51-62,
                 config/structured.py:105-107).  diffusers is not vendored: restated from its published
                 algorithm (epsilon prediction, `fixed_small` variance clamped at 1e-20).  PARITY UNPINNED.
  PVDSchedule    the prior side: the in-repo GaussianDiffusion (reference pvd/__init__.py:18-68
                 coefficients, :136-224 p_mean_variance / p_sample with model_mean_type='eps',
                 model_var_type='fixedsmall', betas linear 1e-4..0.02 `:477`).  Pinned bit for bit against
                 the reference class on CPU (tests/test_reference_pins.py).
  BDMSampler     vanilla PC^2 sampling (model.py:123-214), BDM-Blending (main_blending.py:186-347) and
                 BDM-Merging (main_merging.py:369-523) with the shipped schedule
                 roll_step=16, milestones=[1000,968,936,872,128,64,32,0] as default.

On CUDA a denoising step -- conditioning, denoiser, noise draw, posterior update, timestep decrement -- is ONE
CUDA-graph replay (`GraphedChain`): the timestep lives on the device, the per-timestep coefficients come
from a device table (`bdm_sampler_update`, csrc/sampler.cu), and a chain of k steps is k replays with no
other launch and no host->device traffic.  The eager methods (`DDPMSchedule.step`, `PVDSchedule.step`) are
the same arithmetic as separate torch ops; graph replays are bit-identical to them under the same noise.
"""
import numpy as np
import torch

from . import backend as _backend


class DDPMSchedule:
    def __init__(self, beta_start=1e-5, beta_end=8e-3, num_train_timesteps=1000):
        self.num_train_timesteps = num_train_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.timesteps = list(range(num_train_timesteps - 1, -1, -1))  # set_timesteps(1000)
        self._tables = {}

    def coefficients(self, t):
        """Python floats for one step t -> (sqrt(abar_t), sqrt(1-abar_t), coef_x0, coef_xt, sigma)"""
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

    def row(self, t):
        """(c0..c4) of bdm_sampler_update mode 0 as float32: sqrt(1-abar), 1/sqrt(abar), coef_x0, coef_xt, sigma.
        The reciprocal is taken in double and rounded to float -- what torch's CUDA `tensor / scalar` does."""
        sqrt_a, sqrt_b, coef_x0, coef_xt, sigma = self.coefficients(int(t))
        return np.array([sqrt_b, 1.0 / sqrt_a, coef_x0, coef_xt, sigma, 0, 0, 0], dtype=np.float32)

    def table(self, device):
        key = str(device)
        if key not in self._tables:
            rows = np.stack([self.row(t) for t in range(self.num_train_timesteps)])
            self._tables[key] = torch.from_numpy(rows).to(device)
        return self._tables[key]

    def step(self, eps, t, x_t, generator=None, noise=None):
        """x_{t-1} from the predicted noise (scheduler.step(...).prev_sample), as separate torch ops.  The
        variance noise is drawn on every step (also at t = 0, where it is not used) so that the generator
        advances exactly as it does inside a captured graph."""
        c = self.row(int(t))
        if noise is None:
            noise = torch.randn(eps.shape, generator=generator, device=eps.device, dtype=eps.dtype)
        x0 = (x_t - float(c[0]) * eps) * float(c[1])
        prev = float(c[2]) * x0 + float(c[3]) * x_t
        if int(t) > 0:
            prev = prev + float(c[4]) * noise
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
        self.sigma = torch.exp(0.5 * self.post_log_var)          # p_sample :217, per timestep
        self._tables = {}

    def row(self, t):
        """(c0..c4) of bdm_sampler_update mode 1"""
        t = int(t)
        return np.array([float(self.sqrt_recip_ac[t]), float(self.sqrt_recipm1_ac[t]), float(self.coef1[t]),
                         float(self.coef2[t]), float(self.sigma[t]), 0, 0, 0], dtype=np.float32)

    def table(self, device):
        key = str(device)
        if key not in self._tables:
            rows = np.stack([self.row(t) for t in range(self.num_timesteps)])
            self._tables[key] = torch.from_numpy(rows).to(device)
        return self._tables[key]

    def step(self, eps, t, x_t, generator=None, noise=None):
        t = int(t)
        c = self.row(t)
        x0 = float(c[0]) * x_t - float(c[1]) * eps
        mean = float(c[2]) * x0 + float(c[3]) * x_t
        if noise is None:
            noise = torch.randn(x_t.shape, generator=generator, device=x_t.device, dtype=x_t.dtype)
        if t == 0:
            return mean
        return mean + float(c[4]) * noise


class GraphedStep:
    """CUDA-graph capture of a pure function of two tensors (e.g. `eps = denoiser(x, t)`) with static input /
    output buffers: `g(x, t)` copies the arguments in and replays.  Same kernels on the same data as the eager
    call, so the result is bit-identical.  (The samplers use GraphedChain, which also owns the update.)"""

    def __init__(self, fn, x_example, t_example, warmup=3):
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


class GraphedChain:
    """One denoising step captured as a CUDA graph that advances its own state.

    `body(state)` must read `state.x` (and any extra static inputs), read the timestep from `state.t`
    (int32[1] on the device), write the updated cloud back into `state.x` and decrement `state.t`.
    After capture, `run(x, t_first, steps)` copies the cloud and the first timestep in and replays
    `steps` times: nothing else is launched and nothing crosses PCIe between the steps.  The noise is
    drawn inside the graph from the sampler's generator (registered with the graph, so replays consume
    the generator exactly like eager calls do)."""

    def __init__(self, body, x_example, generator=None, extra=None, warmup=2):
        dev = x_example.device
        self.x = x_example.clone()
        self.t = torch.full((1,), 500, dtype=torch.int32, device=dev)
        self.extra = {k: v.clone() for k, v in (extra or {}).items()}
        self.steps_replayed = 0
        saved_rng = generator.get_state() if generator is not None else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self.t.fill_(500)
                body(self)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.t.fill_(500)
        self.graph = torch.cuda.CUDAGraph()
        if generator is not None:
            self.graph.register_generator_state(generator)
        n0 = _backend.LAUNCHES
        with torch.cuda.graph(self.graph), torch.no_grad():
            body(self)
        self.launches_per_replay = _backend.LAUNCHES - n0     # libbdm_b200 kernels inside one replay
        if generator is not None:
            generator.set_state(saved_rng)                     # warm-up and capture must not consume randomness

    def run(self, x, t_first, steps=1, **extra):
        self.x.copy_(x)
        for k, v in extra.items():
            self.extra[k].copy_(v)
        self.t.fill_(int(t_first))
        for _ in range(steps):
            self.graph.replay()
        self.steps_replayed += steps
        _backend._count_launches(self.launches_per_replay * steps)
        return self.x.clone()


DEFAULT_MILESTONES = (1000, 968, 936, 872, 128, 64, 32, 0)
DEFAULT_ROLL_STEP = 16


class BDMSampler:
    """Holds the co-resident networks of one rank and runs the coupled chains for a batch of shapes.

    pc2_net:  PointCloudModel  ((B,N,3+C), t) -> (B,N,3)
    pvd_net:  PVCNN2_PVD       ((B,3,N), t)   -> (B,3,N)        (optional)
    fuse_net: PVCNNFuse                                          (optional, Merging only)
    conditioner: projection.ProjectionConditioner for the batch

    Captured graphs bake in the addresses of the networks' weights and of the conditioner's tensors:
    assigning a new network or conditioner drops them (re-enable with enable_cuda_graphs); to serve
    another batch of shapes with the same graphs use `conditioner.load(...)`, which refills the
    conditioner's tensors in place.
    """

    def __init__(self, pc2_net, conditioner, pvd_net=None, fuse_net=None, generator=None):
        self._graphs = {}
        self._pc2_net, self._pvd_net, self._fuse_net, self._cond = pc2_net, pvd_net, fuse_net, conditioner
        self.gen = generator
        self.ddpm = DDPMSchedule()
        self.pvd = PVDSchedule()
        self.forwards = dict(pc2=0, pvd=0, fuse=0)

    def _swap(self, name, value):
        if getattr(self, name) is not value:
            self._graphs.clear()      # a replay would silently keep using the old object's memory
        setattr(self, name, value)

    pc2_net = property(lambda self: self._pc2_net, lambda self, v: self._swap("_pc2_net", v))
    pvd_net = property(lambda self: self._pvd_net, lambda self, v: self._swap("_pvd_net", v))
    fuse_net = property(lambda self: self._fuse_net, lambda self, v: self._swap("_fuse_net", v))
    cond = property(lambda self: self._cond, lambda self, v: self._swap("_cond", v))

    # -- noise predictions -----------------------------------------------------------------------
    def _pc2_eps(self, x_t, tt):
        """noise prediction of the PC^2 branch: conditioning + denoiser.  With a channel-last feature map
        and a CUDA conditioner the projected features are written directly in the denoiser's
        channel-first layout (one pass instead of gather + concat + transpose)."""
        fused = getattr(self.cond, "channel_last", False) and hasattr(self.cond, "get_input_channel_first") \
            and hasattr(self.pc2_net, "forward_channel_first") and x_t.is_cuda
        if fused:
            return self.pc2_net.forward_channel_first(self.cond.get_input_channel_first(x_t), tt)
        return self.pc2_net(self.cond.get_input_with_conditioning(x_t), tt)

    def _fuse_eps(self, recon_centred, prior_centred, tt):
        """model.py:533-560: condition the PC^2 branch, run PVCNN_fuse on (conditioned recon, prior)"""
        if getattr(self.cond, "channel_last", False) and hasattr(self.cond, "get_input_channel_first") \
                and recon_centred.is_cuda:
            cond_cf = self.cond.get_input_channel_first(recon_centred)
        else:
            cond_cf = self.cond.get_input_with_conditioning(recon_centred).transpose(1, 2)
        return self.fuse_net(cond_cf, prior_centred.transpose(1, 2).contiguous(), tt).transpose(1, 2)

    # -- CUDA graphs -----------------------------------------------------------------------------
    def enable_cuda_graphs(self, x_example):
        """Capture one whole step of each chain for clouds shaped like `x_example` (B,N,3): the PC^2 step, and
        when the networks are present the PVD step and the Merging fusion step.  Later steps with that
        shape replay the graphs; other shapes run eagerly."""
        b = x_example.shape[0]
        dev = x_example.device
        ddpm_table, pvd_table = self.ddpm.table(dev), self.pvd.table(dev)

        def pc2_body(st):
            tt = st.t.to(torch.long).expand(b)
            eps = self._pc2_eps(st.x, tt).contiguous()
            noise = torch.randn(st.x.shape, generator=self.gen, device=dev, dtype=st.x.dtype)
            _backend.sampler_update(st.x, eps, noise, ddpm_table, st.t, 0, out=st.x)
            st.t.sub_(1)

        self._graphs[("pc2", tuple(x_example.shape))] = GraphedChain(pc2_body, x_example, self.gen)
        if self.pvd_net is not None:
            x_cf = x_example.permute(0, 2, 1).contiguous()

            def pvd_body(st):
                tt = st.t.to(torch.long).expand(b)
                eps = self.pvd_net(st.x, tt).contiguous()
                noise = torch.randn(st.x.shape, generator=self.gen, device=dev, dtype=st.x.dtype)
                _backend.sampler_update(st.x, eps, noise, pvd_table, st.t, 1, out=st.x)
                st.t.sub_(1)

            self._graphs[("pvd", tuple(x_cf.shape))] = GraphedChain(pvd_body, x_cf, self.gen)
        if self.fuse_net is not None:
            def fuse_body(st):
                tt = st.t.to(torch.long).expand(b)
                prior = st.extra["prior"]
                prior_c = prior - prior.mean(dim=1, keepdim=True)
                recon_c = st.x - st.x.mean(dim=1, keepdim=True)
                eps = self._fuse_eps(recon_c, prior_c, tt).contiguous()
                noise = torch.randn(st.x.shape, generator=self.gen, device=dev, dtype=st.x.dtype)
                _backend.sampler_update(recon_c, eps, noise, ddpm_table, st.t, 0, out=st.x)
                st.t.sub_(1)

            self._graphs[("fuse", tuple(x_example.shape))] = GraphedChain(fuse_body, x_example, self.gen,
                                                                           extra={"prior": x_example})

    @property
    def graph_launches_per_step(self):
        """libbdm_b200 kernels inside one replay of each captured graph"""
        return {k[0]: g.launches_per_replay for k, g in self._graphs.items()}

    # -- one denoising step of each kind ---------------------------------------------------------
    def pc2_step(self, x_t, t):
        graphed = self._graphs.get(("pc2", tuple(x_t.shape)))
        self.forwards['pc2'] += 1
        if graphed is not None:
            return graphed.run(x_t, t, 1)
        tt = torch.full((x_t.shape[0],), int(t), device=x_t.device, dtype=torch.long)
        return self.ddpm.step(self._pc2_eps(x_t, tt), t, x_t, self.gen)

    def pvd_step(self, x_t_cf, t):
        """x_t_cf channel-first (B,3,N)"""
        graphed = self._graphs.get(("pvd", tuple(x_t_cf.shape)))
        self.forwards['pvd'] += 1
        if graphed is not None:
            return graphed.run(x_t_cf, t, 1)
        tt = torch.full((x_t_cf.shape[0],), int(t), device=x_t_cf.device, dtype=torch.long)
        return self.pvd.step(self.pvd_net(x_t_cf, tt), t, x_t_cf, self.gen)

    # -- chains ----------------------------------------------------------------------------------
    def pc2_chain(self, x, start_time, end_time):
        """model.py:216-289 interaction_sample: timesteps[1000-start : 1000-end] = start-1 ... end"""
        steps = start_time - end_time
        graphed = self._graphs.get(("pc2", tuple(x.shape)))
        if graphed is not None and steps > 0:
            self.forwards['pc2'] += steps
            return graphed.run(x, start_time - 1, steps)
        for t in range(start_time - 1, end_time - 1, -1):
            x = self.pc2_step(x, t)
        return x

    def pvd_chain(self, x, start_time, final_time):
        """pvd/__init__.py:450-473 generate_pvd_xyz on (B,N,3) clouds (main_blending.py:176-183)"""
        x = x.permute(0, 2, 1).float().contiguous()
        steps = start_time - final_time
        graphed = self._graphs.get(("pvd", tuple(x.shape)))
        if graphed is not None and steps > 0:
            self.forwards['pvd'] += steps
            x = graphed.run(x, start_time - 1, steps)
        else:
            for t in reversed(range(final_time, start_time)):
                x = self.pvd_step(x, t)
        return x.permute(0, 2, 1)

    def fuse_step(self, from_prior, from_recon, timestep):
        """model.py:510-570 nstep_fuse: recentre both clouds (:530-531), one PVCNN_fuse forward, one DDPM step
        from the recentred PC^2 sample (:563-565)"""
        self.forwards['fuse'] += 1
        graphed = self._graphs.get(("fuse", tuple(from_recon.shape)))
        if graphed is not None:
            return graphed.run(from_recon.contiguous(), timestep, 1, prior=from_prior.contiguous())
        from_prior = from_prior - from_prior.mean(dim=1, keepdim=True)
        from_recon = from_recon - from_recon.mean(dim=1, keepdim=True)
        tt = torch.full((from_recon.shape[0],), int(timestep), device=from_recon.device, dtype=torch.long)
        eps = self._fuse_eps(from_recon, from_prior, tt)
        return self.ddpm.step(eps, timestep, from_recon, self.gen)

    def _init_cloud(self, b, n, device, centre):
        x = torch.randn(b, n, 3, generator=self.gen, device=device)
        return x - x.mean(dim=1, keepdim=True) if centre else x

    def _branch_mask(self, b, n, device, mask_generator):
        """main_blending.py:330-344: a fair coin per point.  The reference draws it on the host
        (`torch.randint(0, 2, (B, N))`) and uploads it; with no host generator given it is drawn on the
        device from the sampler's generator (no PCIe copy inside the sampling loop)."""
        if mask_generator is not None:
            return torch.randint(0, 2, (b, n), generator=mask_generator).to(device).bool()
        return torch.randint(0, 2, (b, n), generator=self.gen, device=device).bool()

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
                recon = self.pc2_chain(x, m[i + 1], m[i + 1] - roll_step)
                prior = self.pvd_chain(x, m[i + 1], m[i + 1] - roll_step)
                pick = self._branch_mask(b, n, device, mask_generator)
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
                recon = self.pc2_chain(x, m[i + 1], m[i + 1] - roll_step + 1)
                prior = self.pvd_chain(x, m[i + 1], m[i + 1] - roll_step + 1)
                x = self.fuse_step(prior, recon, m[i + 1] - roll_step)
        return x


def forward_counts(milestones=DEFAULT_MILESTONES, roll_step=DEFAULT_ROLL_STEP, mode="merging"):
    """Denoiser forwards per shape implied by a schedule."""
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
