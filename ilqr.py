"""Drop-in module name of the reference (`from ilqr import IterativeLinearQuadraticRegulator`,
/root/reference/pendulum.py:11): re-exports the B200 solver class."""
from drake_ddp_b200.ilqr import BatchedILQR, IterativeLinearQuadraticRegulator  # noqa: F401
