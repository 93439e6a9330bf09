"""consolver_b200 — B200-native (sm_100a) implementation of the ConsistencySolver sampling step of
G-U-N/consolver, behind the reference's own scheduler plugin API.

    from consolver_b200 import PPOScheduler, FMPPOScheduler     # drop-ins for scheduler_ppo / scheduler_fmppo
    from consolver_b200 import FlowMatchGeneralDiscreteScheduler  # the baselines of edit_ppo/scheduler_fm
    from consolver_b200 import DPMSolverMultistepScheduler        # diffusers_amed_plugin_dpmpp (AMED baseline)

Host side: Python/PyTorch (device memory, streams, RNG, torch.distributed); hot path: hand-written CUDA behind
the C ABI in include/consolver.h (libconsolver.so, built in-tree by consolver_b200.build).  No CPU fallback."""
from .factor_net import FactorNetPPO, FactorNetPPOContinous, FactorNetPPOFM
from .scheduler_dpm import DPMSolverMultistepScheduler
from .scheduler_fm import FlowMatchGeneralDiscreteScheduler
from .scheduler_fmppo import FMPPOScheduler, FMPPOSchedulerOutput
from .scheduler_ppo import PPOScheduler, PPOSchedulerOutput

__all__ = ["PPOScheduler", "PPOSchedulerOutput", "FMPPOScheduler", "FMPPOSchedulerOutput", "FactorNetPPO",
           "FactorNetPPOFM", "FactorNetPPOContinous", "FlowMatchGeneralDiscreteScheduler", "DPMSolverMultistepScheduler"]
__version__ = "0.2.0"
