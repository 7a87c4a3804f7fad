from .models import MLP, MLP_Embedding  # noqa: F401
