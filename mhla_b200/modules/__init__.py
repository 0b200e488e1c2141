"""Drop-in ``nn.Module`` shims with the reference's constructor signatures and ``state_dict`` keys
(SURVEY.md 8b); the token-mixing operator inside each of them is the CUDA kernel."""
from .dit import MHLA4DiT, MHLA_Normed_Torch  # noqa: F401
from .wan import MHLA_Video_Uni, WanRMSNorm, rope_apply  # noqa: F401
from .nlp import MHLA  # noqa: F401
