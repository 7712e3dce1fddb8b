"""Method configurations of the reference (/root/reference/dn_splatter/dn_config.py:13-198) as plain data:
the three method names, their model-config overrides and the per-parameter-group Adam settings.  When
nerfstudio is importable, `method_specifications()` wraps them into MethodSpecification objects for the
same entry points (pyproject.toml:27-42 of the reference); without it the dictionaries drive
dn_pipeline.DNSplatterPipeline / the bundled trainer loop directly."""
from __future__ import annotations

from typing import Dict

from .dn_model import DNSplatterModelConfig

MAX_NUM_ITERATIONS = 30000  # dn_config.py:20


def optimizer_groups(max_steps: int = MAX_NUM_ITERATIONS) -> Dict[str, Dict]:
    """lr / eps / exponential-decay target per gauss_params group (dn_config.py:29-68)."""
    g = lambda lr, final=None: {"lr": lr, "eps": 1e-15, "lr_final": final, "max_steps": max_steps}  # noqa: E731
    return {
        "means": g(1.6e-4, 1.6e-6), "features_dc": g(0.0025), "features_rest": g(0.0025 / 20), "opacities": g(0.05),
        "scales": g(0.005), "quats": g(0.001), "camera_opt": g(1e-3, 5e-5),
        "normals": g(1e-3),  # no-op group: "its just here to make the trainer happy" (dn_config.py:62-67)
    }


METHODS: Dict[str, Dict] = {
    "dn-splatter": {
        "description": "DN-Splatter: depth and normal priors for 3DGS",
        "model": lambda: DNSplatterModelConfig(regularization_strategy="dn-splatter"),
    },
    "ags-mesh": {
        "description": "AGS-Mesh: adaptive Gaussian splatting and meshing",
        "model": lambda: DNSplatterModelConfig(regularization_strategy="ags-mesh"),
    },
    "dn-splatter-big": {  # dn_config.py:137-198 (:150-153): lower cull threshold, no culling after stop_split_at
        "description": "DN-Splatter Big variant",
        "model": lambda: DNSplatterModelConfig(regularization_strategy="dn-splatter", cull_alpha_thresh=0.005,
                                               continue_cull_post_densification=False),
    },
}
TRAINER_DEFAULTS = dict(steps_per_eval_image=500, steps_per_eval_batch=500, steps_per_save=1000000,
                        steps_per_eval_all_images=1000000, max_num_iterations=MAX_NUM_ITERATIONS, mixed_precision=False,
                        gradient_accumulation_steps={"camera_opt": 100, "color": 10, "shs": 10})


def method_specifications():
    """nerfstudio MethodSpecification objects (only when nerfstudio is installed)."""
    from nerfstudio.plugins.types import MethodSpecification  # noqa: F401  (raises ImportError otherwise)

    raise NotImplementedError("nerfstudio is not part of this image; wire METHODS into TrainerConfig where it exists "
                              "(INTEGRATION.md shows the three-line registration)")
