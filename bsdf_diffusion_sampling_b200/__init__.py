"""bsdf_diffusion_sampling_b200 -- B200-native (sm_100a) neural BSDF sampler.

Drop-in for the hot path of fzy28/BSDF_diffusion_sampling: the reference's sampler functions
(``mlp_brdf_sampling``), model containers (``model``), the tensor part of its three Mitsuba BSDF
plugins (``plugins``), the reflow ``dosampling`` loop and the tinycudann ``Network`` shim (``reflow``),
all backed by one C-ABI CUDA library (``libbsdfdiff.so``, ``include/bsdfdiff.h``).

Importing the package requires the built library; there is no CPU or PyTorch fallback.
"""
import sys as _sys

__all__ = []
if "bsdf_diffusion_sampling_b200.build" not in getattr(_sys, "orig_argv", []):
    # (`python -m bsdf_diffusion_sampling_b200.build` must be able to run when the library is missing or stale)
    from . import _lib  # noqa: F401  (raises ImportError if libbsdfdiff.so is missing)
    from . import ops, weights, model, mlp_brdf_sampling, reflow, measured, plugins, sharding, training, materials, mcmc  # noqa: F401
    from .mlp_brdf_sampling import (  # noqa: F401
        network_sampling_disk, network_sampling_disk_tiny, network_pdf_disk,
        network_sampling_spherical, network_pdf_spherical,
    )
    from .ops import set_default_precision  # noqa: F401

    __all__ = [
        "ops", "weights", "model", "mlp_brdf_sampling", "reflow", "plugins", "sharding", "training", "materials", "mcmc",
        "network_sampling_disk", "network_sampling_disk_tiny", "network_pdf_disk",
        "network_sampling_spherical", "network_pdf_spherical", "set_default_precision",
    ]
