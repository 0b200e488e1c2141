"""Drop-in ``nn.Module`` shims with the reference's constructor signatures and ``state_dict`` keys
(SURVEY.md 8b); the token-mixing operator inside each of them is the CUDA kernel."""
from .dit import MHLA4DiT, MHLA_Normed_Torch  # noqa: F401
from .wan import (  # noqa: F401
    MHLA_Video_Uni, Gated_MHLA_Video, MHLA_Video_Nope, Gated_MHLA_Video_LePE, MHLA_Video_LePE, MHLA_Video,
    WAN_SELFATTENTION_CLASSES, WanRMSNorm, rope_apply,
)
from .nlp import MHLA  # noqa: F401
