"""Shared policy network of the RL shell on the CUDA path.

``SharedMLPPolicy`` holds the weights of the reference's ``CustomModel1`` (run_rllib.py:35-52:
``fc1`` 64->64 ReLU, ``fc2`` 64->64 ReLU, ``fc_out`` 64->num_outputs) as torch CUDA tensors and
evaluates it for every agent of a batch with one launch of ``policy_mlp_kernel``
(``orca_policy_mlp`` in include/orca_b200.h).  With it the loop

    obs = env.reset()
    while True:
        obs, reward, done, _ = env.step(policy.act(obs))

runs without a host round trip.  Training (PPO in the reference, through RLlib) is out of scope;
weights come from ``load_state`` (e.g. exported from a trained model) or random init.

torch is plumbing (device memory, streams); there is no torch fallback for the forward pass.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, Optional

import torch

from . import _lib

IN_DIM = 64      # Collision_Avoidance_Env observation: 16 rays x (hit.x, hit.y, vel.x, vel.y)
HIDDEN_DIM = 64  # run_rllib.py:45,47


class SharedMLPPolicy:
    """``CustomModel1`` (run_rllib.py:35-52) for a batch of agents.

    ``num_outputs`` = 2 for PPO's diagonal Gaussian over the env's Box(1) action
    (collision_avoidence_env.py:52-53): column 0 is the mean heading change, column 1 the
    log-std.  Weight matrices are stored [in][out] like ``slim.fully_connected``.
    """

    def __init__(self, sim, num_outputs: int = 2, seed: int = 0, impl: str = "tcgen05"):
        if not 1 <= num_outputs <= 8:
            raise NotImplementedError("num_outputs must be in [1, 8]")
        if impl not in ("tcgen05", "fp32"):
            raise ValueError("impl must be 'tcgen05' (tensor cores, 3xTF32) or 'fp32' (FP32 pipes)")
        self.impl = impl
        self._sim = sim
        self._L = _lib.load()
        self.device = sim.device
        self.num_outputs = int(num_outputs)
        g = torch.Generator().manual_seed(seed)

        def xavier(fan_in, fan_out):  # slim.fully_connected default initializer (xavier uniform)
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            return ((torch.rand(fan_in, fan_out, generator=g) * 2 - 1) * lim).to(self.device)

        self.w1, self.b1 = xavier(IN_DIM, HIDDEN_DIM), torch.zeros(HIDDEN_DIM, device=self.device)
        self.w2, self.b2 = xavier(HIDDEN_DIM, HIDDEN_DIM), torch.zeros(HIDDEN_DIM, device=self.device)
        self.w3, self.b3 = xavier(HIDDEN_DIM, self.num_outputs), torch.zeros(self.num_outputs, device=self.device)
        self._out: Optional[torch.Tensor] = None

    # -------------------------------------------------------------------- weights
    def state(self) -> Dict[str, torch.Tensor]:
        return {"fc1/weights": self.w1, "fc1/biases": self.b1, "fc2/weights": self.w2, "fc2/biases": self.b2,
                "fc_out/weights": self.w3, "fc_out/biases": self.b3}

    def load_state(self, state: Dict[str, torch.Tensor]) -> None:
        """Weights keyed like the reference's TF variables (``fc1/weights`` [64, 64] ...)."""
        for name, dst in self.state().items():
            src = torch.as_tensor(state[name], dtype=torch.float32)
            if tuple(src.shape) != tuple(dst.shape):
                raise ValueError(f"{name}: expected shape {tuple(dst.shape)}, got {tuple(src.shape)}")
            dst.copy_(src)

    # -------------------------------------------------------------------- forward
    def forward(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``obs`` [..., 64] float32 CUDA (contiguous) -> [..., num_outputs]."""
        if obs.device.type != "cuda" or obs.dtype != torch.float32 or not obs.is_contiguous():
            raise ValueError("obs must be a contiguous float32 CUDA tensor")
        if obs.shape[-1] != IN_DIM:
            raise ValueError(f"obs rows must have {IN_DIM} floats")
        rows = obs.numel() // IN_DIM
        shape = tuple(obs.shape[:-1]) + (self.num_outputs,)
        if out is None:
            if self._out is None or tuple(self._out.shape) != shape:
                self._out = torch.empty(shape, dtype=torch.float32, device=obs.device)
            out = self._out
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != obs.device:
            raise ValueError("out must be a contiguous float32 tensor of shape " + str(shape))
        w = _lib.OrcaMlpWeights(ctypes.sizeof(_lib.OrcaMlpWeights), IN_DIM, HIDDEN_DIM, self.num_outputs,
                                self.w1.data_ptr(), self.b1.data_ptr(), self.w2.data_ptr(), self.b2.data_ptr(),
                                self.w3.data_ptr(), self.b3.data_ptr())
        stream = torch.cuda.current_stream(obs.device).cuda_stream
        fn = self._L.orca_policy_mlp if self.impl == "tcgen05" else self._L.orca_policy_mlp_fp32
        _lib.check(fn(self._sim._h, obs.data_ptr(), rows, ctypes.byref(w), out.data_ptr(), ctypes.c_void_p(stream)))
        return out

    __call__ = forward

    def act(self, obs: torch.Tensor, deterministic: bool = True, generator: Optional[torch.Generator] = None,
            clip: float = math.pi) -> torch.Tensor:
        """Heading change per agent ([...] float32): the Gaussian mean (column 0), or a sample with
        the log-std of column 1; clipped to the env's action box."""
        y = self.forward(obs)
        mean = y[..., 0]
        if not deterministic and self.num_outputs >= 2:
            noise = torch.randn(mean.shape, device=mean.device, generator=generator)
            mean = mean + torch.exp(y[..., 1]) * noise
        return mean.clamp(-clip, clip)
