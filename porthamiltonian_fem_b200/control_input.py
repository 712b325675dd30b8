"""Host-side mirror of the reference's passivity controller (reference control_input.py:3-20).

In the reference the controller is a dolfin ``UserExpression`` that Python evaluates at every quadrature point of the
port boundary.  Here the law is a scalar evaluated on the device inside the step loop (``input_value`` in
csrc/phwave.cu, PHW_IN_PASSIVITY).  This module keeps the reference's class name and call surface for host-side
checks and for user code that constructs the controller object; the law itself lives in ``passivity_input``."""
import math

EXCITATION_AMPLITUDE = 10.0          # u(t) = 10 sin(8 pi t) while t < input_stop_t
EXCITATION_OMEGA = 8.0 * math.pi
DAMPING_GAIN = 100.0                 # u = -100 h_1 once t >= control_start_t (h_1: previous step's first lumped state)


def passivity_input(t, h_1, input_stop_t, control_start_t):
    """Three regimes in the reference's order (control_input.py:12-18; same constants and order as `input_value`,
    PHW_IN_PASSIVITY, in csrc/phwave.cu): excitation, free evolution, damping injection."""
    if t < input_stop_t:
        return EXCITATION_AMPLITUDE * math.sin(EXCITATION_OMEGA * t)
    elif t < control_start_t:
        return 0.0
    else:
        return -DAMPING_GAIN * h_1


class Passive_control_input:
    """Carries (t, h_1, input_stop_t, control_start_t) like the reference object; ``t`` and ``h_1`` are updated by the caller."""
    __slots__ = ('t', 'h_1', 'input_stop_t', 'control_start_t', 'eps')

    def __init__(self, t, h_1, input_stop_t, control_start_t, **_ignored):
        self.t, self.h_1 = t, h_1
        self.input_stop_t, self.control_start_t = input_stop_t, control_start_t
        self.eps = 1e-3

    def value(self):
        return passivity_input(self.t, self.h_1, self.input_stop_t, self.control_start_t)

    def eval_cell(self, value, x=None, ufc_cell=None):
        value[0] = self.value()

    @staticmethod
    def value_shape():
        return ()
