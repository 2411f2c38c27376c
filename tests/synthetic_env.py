"""The synthetic Environment lives in the package (simple_dqn_b200/synthetic_env.py); tests import it from here."""
from simple_dqn_b200.synthetic_env import SyntheticEnvironment  # noqa: F401
