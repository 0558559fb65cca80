"""``import enero_b200.actorPPOmiddR as actor`` -- drop-in for the reference's actorPPOmiddR.py:
``actor.myModel(hparams, hidden_init_actor, kernel_init_actor)`` (train_Enero_3top_script.py:119)."""
from .models import ActorModel as myModel  # noqa: F401
