"""Host-side mirror of the reference's passivity controller (control_input.py:3-20).

In the reference this is a dolfin ``UserExpression`` evaluated in Python at every quadrature point of the
port boundary.  Here the controller is a scalar function evaluated on the device inside the step loop
(``input_value`` in csrc/phwave.cu, PHW_IN_PASSIVITY); this class only carries the parameters and offers
the same ``eval`` law for host-side checks."""
import math


class Passive_control_input(object):
    def __init__(self, t, h_1, input_stop_t, control_start_t, **kwargs):
        self.t = t
        self.h_1 = h_1
        self.input_stop_t = input_stop_t
        self.control_start_t = control_start_t
        self.eps = 0.001

    def value(self):
        if self.t < self.input_stop_t:
            return 10 * math.sin(8 * math.pi * self.t)
        elif self.t < self.control_start_t:
            return 0.0
        return -100.0 * self.h_1

    def eval_cell(self, value, x, ufc_cell):
        value[0] = self.value()

    def value_shape(self):
        return ()
