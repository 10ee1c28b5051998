"""instant_angelo_b200 -- B200-native (sm_100a) implementation of Instant-angelo's training hot path.

Host-side mirror of the reference's Python surface:
  network_utils  <- models/network_utils.py   (get_encoding, get_mlp, ProgressiveBandHashGrid, VanillaMLP, ...)
  geometry       <- models/geometry.py        (VolumeSDF, VolumeDensity, contract_to_unisphere)
  texture        <- models/texture.py         (VolumeRadiance, VolumeDualColor, VolumeDualColorV3)
  neus           <- models/neus.py            (VarianceNetwork, NeuSModel)
  nerfacc_api    <- nerfacc==0.3.3            (OccupancyGrid, ray_marching, render_weight_from_alpha, ...)
  losses         <- systems/neus.py:130-194   (training_loss), systems/base.py:28-45 (C)
  systems        <- systems/neus.py, systems/base.py, systems/utils.py:314-346, models/ray_utils.py
                    (NeuSSystem.preprocess_data / training_step / validation_step, parse_optimizer, parse_scheduler, get_rays)
All arithmetic runs in csrc/ (CUDA, C ABI in include/ia_b200.h).  There is no CPU fallback.
"""
from . import registry as models  # noqa: F401
from .registry import make, register  # noqa: F401

__all__ = ["models", "make", "register"]

from . import neus as _neus  # noqa: E402,F401  (registers 'neus', 'volume-sdf', 'volume-density' and the colour heads)
