"""Functional (explicit-state, ask/tell) counterparts of the distribution-based searchers and their optimizers
(reference: evotorch/algorithms/functional/__init__.py).  Every function accepts extra leftmost batch dimensions.

    state = pgpe(center_init=x0, center_learning_rate=0.5, stdev_learning_rate=0.1, stdev_init=1.0, objective_sense="min")
    for _ in range(generations):
        population = pgpe_ask(state, popsize=1000)
        state = pgpe_tell(state, population, f(population))
    best_guess = state.optimizer_state.center
"""

from .funcadam import AdamState, adam, adam_ask, adam_tell
from .funccem import CEMState, cem, cem_ask, cem_tell
from .funcclipup import ClipUpState, clipup, clipup_ask, clipup_tell
from .funcpgpe import PGPEState, pgpe, pgpe_ask, pgpe_tell
from .funcsgd import SGDState, sgd, sgd_ask, sgd_tell
from .misc import OptimizerFunctions, get_functional_optimizer

__all__ = ["AdamState", "adam", "adam_ask", "adam_tell", "CEMState", "cem", "cem_ask", "cem_tell", "ClipUpState", "clipup", "clipup_ask",
           "clipup_tell", "PGPEState", "pgpe", "pgpe_ask", "pgpe_tell", "SGDState", "sgd", "sgd_ask", "sgd_tell", "OptimizerFunctions",
           "get_functional_optimizer"]
