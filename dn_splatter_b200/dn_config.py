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
        "description": "AGS-Mesh variant of Splatfacto model. Incorporates depth and normal filtering strategy.",
        "model": lambda: DNSplatterModelConfig(regularization_strategy="ags-mesh"),
    },
    "dn-splatter-big": {  # dn_config.py:137-198 (:150-153): lower cull threshold, no culling after stop_split_at
        "description": "DN-Splatter Big variant",
        "model": lambda: DNSplatterModelConfig(cull_alpha_thresh=0.005, continue_cull_post_densification=False),
    },
}
TRAINER_DEFAULTS = dict(steps_per_eval_image=500, steps_per_eval_batch=500, steps_per_save=1000000,
                        steps_per_eval_all_images=1000000, max_num_iterations=MAX_NUM_ITERATIONS, mixed_precision=False,
                        gradient_accumulation_steps={"camera_opt": 100, "color": 10, "shs": 10})


def method_specifications(datamanager_config=None):
    """The three nerfstudio MethodSpecification objects of the reference's dn_config.py (`dn_splatter`, `ags_mesh`,
    `dn_splatter_big`; entry points of its pyproject.toml:27-42), with THIS package's pipeline and model configs.

    Needs nerfstudio (absent from this image: ImportError).  `datamanager_config`: the datamanager config every method
    uses; default = the reference's own data stack when it is installed next to nerfstudio
    (DNSplatterManagerConfig(dataparser=NormalNerfstudioConfig(load_3D_points=True)), dn_config.py:24-27) — data parsing is
    outside this package's scope (SURVEY §2 out-of-scope rows).  Register in the host package's pyproject.toml:

        [project.entry-points.'nerfstudio.method_configs']
        dn-splatter = 'dn_splatter_b200.dn_config:dn_splatter'
    """
    from nerfstudio.configs.base_config import ViewerConfig
    from nerfstudio.engine.optimizers import AdamOptimizerConfig
    from nerfstudio.engine.schedulers import ExponentialDecaySchedulerConfig
    from nerfstudio.engine.trainer import TrainerConfig
    from nerfstudio.plugins.types import MethodSpecification

    from .dn_pipeline import DNSplatterPipelineConfig

    def datamanager():
        if datamanager_config is not None:
            import copy

            return copy.deepcopy(datamanager_config)
        from dn_splatter.data.normal_nerfstudio import NormalNerfstudioConfig  # the reference's data stack
        from dn_splatter.dn_datamanager import DNSplatterManagerConfig

        return DNSplatterManagerConfig(dataparser=NormalNerfstudioConfig(load_3D_points=True))

    def optimizers():
        out = {}
        for name, g in optimizer_groups().items():
            sched = (ExponentialDecaySchedulerConfig(lr_final=g["lr_final"], max_steps=g["max_steps"])
                     if g["lr_final"] is not None else None)
            out[name] = {"optimizer": AdamOptimizerConfig(lr=g["lr"], eps=g["eps"]), "scheduler": sched}
        return out

    specs = {}
    for name, m in METHODS.items():
        trainer_kw = dict(TRAINER_DEFAULTS)
        if name == "dn-splatter-big":  # the reference's big variant sets no gradient accumulation (dn_config.py:137-146)
            trainer_kw.pop("gradient_accumulation_steps")
        specs[name] = MethodSpecification(
            config=TrainerConfig(method_name=name, pipeline=DNSplatterPipelineConfig(datamanager=datamanager(), model=m["model"]()),
                                 optimizers=optimizers(), viewer=ViewerConfig(num_rays_per_chunk=1 << 15), vis="viewer",
                                 **trainer_kw),
            description=m["description"])
    return specs


def __getattr__(name):
    """`dn_splatter`, `ags_mesh`, `dn_splatter_big`: module attributes for nerfstudio's entry points, built lazily so that
    importing this module never requires nerfstudio."""
    key = {"dn_splatter": "dn-splatter", "ags_mesh": "ags-mesh", "dn_splatter_big": "dn-splatter-big"}.get(name)
    if key is None:
        raise AttributeError(name)
    return method_specifications()[key]
