"""`simulator._target_` for CPU tests: the product's B200Sim plugin, unchanged, with the CPU oracle standing in for the
CUDA engine behind its two-method backend shape (sim_step, body_states).  TEST INFRASTRUCTURE."""
import numpy as np
import torch

from spi_active_b200 import go2_model as gm
from spi_active_b200.simulator import B200Sim


class OracleBackend:
    def __init__(self, blob):
        self.blob = blob

    def sim_step(self, state, torques, n_steps=1, params=None, param_names=(), flags=0, foot_force=None, ext_wrench=None):
        from oracle import oracle as orc
        ids = [gm.PARAM_IDS[n] for n in param_names]
        new, ff = orc.sim_step(self.blob, state.numpy().astype(np.float64), torques.numpy().astype(np.float64), n_steps,
                               params=None if params is None else params.numpy(), param_ids=ids, flags=flags,
                               return_foot_force=True, ext_wrench=None if ext_wrench is None else ext_wrench.numpy())
        state.copy_(torch.from_numpy(new.astype(np.float32)))
        if foot_force is not None:
            foot_force.copy_(torch.from_numpy(ff.astype(np.float32)))
        return state

    def body_states(self, state, out=None):
        from oracle import oracle as orc
        bs = torch.from_numpy(orc.body_states(self.blob, state.numpy()).astype(np.float32))
        if out is None:
            return bs
        out.copy_(bs)
        return out

    def close(self):
        pass


class B200SimOracleBackend(B200Sim):
    def __init__(self, config=None, device="cpu"):
        model = gm.go2_nominal()
        super().__init__(config=config, device="cpu", backend=OracleBackend(gm.build_model_blob(model)), model=model)
