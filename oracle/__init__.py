"""CPU oracle for the DART tray-tilt NMPC hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in float64 numpy, the arithmetic of the reference's
three high-level MPC controllers (all citations relative to the reference
tree):

* PMPC  ``PMPC/src/controller/mpc_3d.py:28-138``
* RMPC  ``RMPC/dev_dual/controller/np_mpc_adaptive_with_linear_regressor.py:10-222``
        and the caller glue ``RMPC/dev_dual/rob_ctrl.py:281-352``
* LMPC  ``LMPC/src/controller/rlmpc2.py:33-80, 236-491, 606-616, 641-668, 742-759``

It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it.  The
product package never does, and fails loudly when its CUDA library is absent.

PARITY UNPINNED.  The arithmetic that produces the reference's numbers lives
in CasADi (``casadi>=3.5.0``, un-pinned, ``PMPC/requirements.txt:10``) and its
bundled IPOPT/MUMPS, neither of which is vendored, installed here, or
installable (no network).  The reference ships no tests, golden vectors or
recorded traces for this path.  The oracle therefore anchors on the NLPs
themselves: every function below is a literal restatement of the reference's
model/cost/constraint code, the NLP is solved to a KKT point by a dense
primal-dual interior-point method (``oracle.ipm``) that shares no code and no
derivative formulas with the CUDA solver (Jacobians here come from complex-step
differentiation of the literal model code), and the optimum is cross-checked
with scipy's SLSQP on a condensed single-shooting form (``oracle.crosscheck``).
Golden vectors under ``tests/golden`` are produced by ``tests/golden/make_golden.py``
from this oracle.

``oracle.ppo`` (the PPO training block of rlmpc2.py) is different: it is restated with torch, the reference's own
library, and pinned to an artefact of the reference -- the policy it trained and checkpointed
(``tests/golden/ppo_reference_checkpoint.npz``, written by ``tests/golden/make_ppo_golden.py``).  It is imported on demand
(it pulls in torch), not with the package.
"""

from . import models, problems, ipm, rls, policy, crosscheck, closed_loop  # noqa: F401
