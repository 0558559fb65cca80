"""``import enero_b200.criticPPO as critic`` -- drop-in for the reference's criticPPO.py:
``critic.myModel(hparams, hidden_init_critic, kernel_init_critic)`` (train_Enero_3top_script.py:122)."""
from .models import CriticModel as myModel  # noqa: F401
