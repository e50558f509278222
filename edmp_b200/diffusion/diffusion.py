"""Diffusion -- reference-shaped boundary class for the guided sampler.

Same constructor and ``denoise_guided`` / ``denoise`` signatures as reference
diffusion/diffusion.py:8-356; the 255-step loop itself (UNet, posterior update, guide gradient,
guided update, endpoint conditioning) runs on the device in one call of edmp_sample_guided with
no host round trips between steps.
"""
import ctypes

import numpy as np
import torch

from .. import _lib


class Diffusion:

    #: how the per-step noise is produced: "numpy" draws x_T and every z from
    #: np.random.multivariate_normal in the reference's order (same stream as the reference under
    #: np.random.seed); "philox" draws z on the device (throughput mode).
    noise_mode = "numpy"
    #: reverse steps per host-noise upload (numpy / recorded-tape modes)
    NOISE_CHUNK = 16

    def __init__(self, T, device, variance_thresh=0.02):
        self.T = T
        self.beta = self.schedule_variance(variance_thresh)
        self.alpha = 1 - self.beta
        self.alpha_bar = np.array([np.prod(self.alpha[:t]) for t in range(1, self.T + 1)])
        self.device = device
        self.variance_thresh = variance_thresh
        self._handle = None
        self._max_rows = 0
        self.last_final_cost = None
        self.last_launches = 0

    def schedule_variance(self, thresh=0.02):
        return np.linspace(0, thresh, self.T + 1)[1:]

    def clip_joints(self, joints):
        lo = np.array([-166, -101, -166, -176, -166, -1, -166]) * (np.pi / 180)
        hi = np.array([166, 101, 166, -4, 166, 215, 166]) * (np.pi / 180)
        return np.clip(joints, lo[np.newaxis, :, np.newaxis], hi[np.newaxis, :, np.newaxis])

    # ---- engine --------------------------------------------------------------------------------------
    def _sampler(self, rows):
        lib = _lib.load()
        if self._handle is None or rows > self._max_rows:
            if self._handle is not None:
                lib.edmp_sampler_destroy(self._handle)
            handle = ctypes.c_void_p()
            _lib.check(lib.edmp_sampler_create(int(self.T), float(self.variance_thresh), int(rows),
                                               ctypes.byref(handle)), "edmp_sampler_create")
            self._handle, self._max_rows = handle, int(rows)
        return self._handle

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().edmp_sampler_destroy(self._handle)
        except Exception:
            pass

    def _draw(self, traj_len, size):
        return np.random.multivariate_normal(mean=np.zeros(traj_len), cov=np.eye(traj_len), size=size)

    def run_steps(self, model, guide, x, start, goal, t_start, t_stop, noise=None, seed=0,
                  guidance_schedule=None, ensemble_rows=None, want_cost=False, condition=True):
        """Device loop over steps t_start .. t_stop+1 on x ([B,7,50] float64 CUDA tensor, in place).
        ``noise``: float64 CUDA tensor [t_start-t_stop, B, 7, 50] or None (Philox with ``seed``)."""
        dev = _lib.require_cuda(self.device)
        lib = _lib.load()
        B = x.shape[0]
        with torch.cuda.device(dev):
            sampler = self._sampler(B)
            unet = model.engine(B)
            scene = None
            if guide is not None:
                scene = guide.scene_handle(rows=B, guidance_schedule=guidance_schedule,
                                           ensemble_rows=ensemble_rows)
            s_arr, s_ptr = _lib.host_f64(start)
            g_arr, g_ptr = _lib.host_f64(goal)
            cost = torch.empty(B, device=dev, dtype=torch.float32) if (want_cost and guide is not None) else None
            _lib.check(lib.edmp_sampler_set_condition(sampler, 1 if condition else 0), "edmp_sampler_set_condition")
            _lib.check(lib.edmp_sample_guided(
                sampler, unet, scene, ctypes.c_void_p(x.data_ptr()), s_ptr, g_ptr,
                ctypes.c_void_p(noise.data_ptr()) if noise is not None else None,
                ctypes.c_uint64(int(seed) & 0xFFFFFFFFFFFFFFFF), B, int(t_start), int(t_stop),
                ctypes.c_void_p(cost.data_ptr()) if cost is not None else None, _lib.stream_ptr()),
                "edmp_sample_guided")
            self.last_launches = int(lib.edmp_sampler_last_launches(sampler))
        return cost

    # ---- reference API -------------------------------------------------------------------------------
    def denoise_guided(self, model, guide, traj_len, num_channels, guidance_schedule, batch_size=1, start=None,
                       goal=None, condition=True, benchmarking=False, noise=None, seed=None):
        """Returns np.float64 [batch_size, 7, 50] like diffusion.py:300-356.  ``noise`` (optional):
        "numpy" | "philox" | (x_T, [z_255 .. z_1]) to replay recorded draws."""
        if traj_len != 50 or num_channels != 7:
            raise ValueError("the engine is compiled for 50 waypoints x 7 joints (cfg1.yaml:16-17)")
        dev = _lib.require_cuda(self.device)
        B = int(batch_size)
        mode = noise if noise is not None else self.noise_mode
        zs = None
        if isinstance(mode, (tuple, list)):
            x_T, zs = mode
            x_T = np.asarray(x_T, dtype=np.float64)
        elif mode in ("numpy", "philox"):
            x_T = self._draw(traj_len, (B, num_channels))
        else:
            raise ValueError("unknown noise mode %r" % (mode,))
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1)) if mode == "philox" else 0
        model.train(False)
        x = torch.as_tensor(x_T, dtype=torch.float64).to(dev).contiguous()
        start, goal = np.asarray(start)[:], np.asarray(goal)[:]
        if mode == "philox":
            cost = self.run_steps(model, guide, x, start, goal, self.T, 0, noise=None, seed=seed,
                                  guidance_schedule=guidance_schedule, want_cost=True, condition=condition)
        else:
            # Host noise (the reference's np.random stream, or a recorded tape) reaches the device in chunks of
            # NOISE_CHUNK steps: the draws of the next chunk overlap the kernels of the current one, and an 8190-row
            # pass never holds more than 16 x 23 MB of noise instead of the whole 5.9 GB tape.
            cost, t = None, self.T
            while t > 0:
                n = min(self.NOISE_CHUNK, t)
                if zs is not None:
                    chunk = np.stack([np.asarray(z, dtype=np.float64) for z in zs[self.T - t:self.T - t + n]])
                else:
                    chunk = np.stack([self._draw(traj_len, (B, num_channels)) for _ in range(n)])
                z = torch.as_tensor(chunk).to(dev).contiguous()
                cost = self.run_steps(model, guide, x, start, goal, t, t - n, noise=z, seed=seed,
                                      guidance_schedule=guidance_schedule, want_cost=(t - n == 0), condition=condition)
                t -= n
        self.last_final_cost = cost
        out = x.cpu().numpy()
        model.check_range()
        return out.copy()

    def denoise(self, model, traj_len, num_channels, start=None, goal=None, condition=True):
        """Unguided sampling of one trajectory (diffusion.py:253-278) -> [7, 50]."""
        dev = _lib.require_cuda(self.device)
        x_T = self._draw(traj_len, (1, num_channels))
        tape = np.stack([self._draw(traj_len, (1, num_channels)) for _ in range(self.T)])
        x = torch.as_tensor(x_T, dtype=torch.float64).to(dev).contiguous()
        z = torch.as_tensor(tape).to(dev).contiguous()
        model.train(False)
        if start is None or goal is None:
            if condition:
                raise ValueError("condition=True needs start and goal")
            start, goal = np.zeros(num_channels), np.zeros(num_channels)   # unused without conditioning and a guide
        self.run_steps(model, None, x, start, goal, self.T, 0, noise=z, condition=condition)
        return x.cpu().numpy()[0]
