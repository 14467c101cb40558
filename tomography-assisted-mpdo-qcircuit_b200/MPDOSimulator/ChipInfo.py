"""Chip presets feeding the noise rates (reference MPDOSimulator/ChipInfo.py:44-87). Times in ns."""

_PRESETS = {
    #          gateTime bath   decay  dephase T1     T2     depolarizing
    'best':   (30,      0.,    0.0,   0.0,    2e11,  2e10,  11e-4),
    'medium': (1,       0.01,  0.98,  0.02,   2e11,  2e10,  5e-2),
    'worst':  (30,      0.,    0.0,   0.0,    2e2,   2e1,   11e-2),
}


class ChipInformation:
    def __init__(self, query_time: str = None):
        self.queryTime = query_time
        self.status = None
        self.gateTime = self.bath_rate = self.dephasing_rate = self.decay_rate = None
        self.T1 = self.T2 = self.chipName = self.dpc_errorRate = None
        self.timeUnit = 'ns'

    def __getattr__(self, item):
        # only reached when normal lookup fails
        raise AttributeError(f'Chip: {item} is not supported.')

    def configure_chip(self, name: str, gate_time: float, decay_rate: float, dephasing_rate: float,
                       bath_rate: float, T1: float, T2: float, dpc_error_rate: float, status: bool):
        self.chipName, self.gateTime, self.bath_rate = name, gate_time, bath_rate
        self.decay_rate, self.dephasing_rate = decay_rate, dephasing_rate
        self.T1, self.T2, self.dpc_errorRate, self.status = T1, T2, dpc_error_rate, status

    def _preset(self, name):
        if self.queryTime is None:
            gt, bath, dec, deph, t1, t2, dpc = _PRESETS[name]
            self.configure_chip(name=name, gate_time=gt, bath_rate=bath, decay_rate=dec, dephasing_rate=deph,
                                T1=t1, T2=t2, dpc_error_rate=dpc, status=True)
        return self

    def best(self):
        return self._preset('best')

    def medium(self):
        return self._preset('medium')

    def worst(self):
        return self._preset('worst')

    def show_property(self):
        print(f"The chip name is: {self.chipName}")
        print(f"The gate time is: {self.gateTime} {self.timeUnit}")
        print(f"The T1 time is: {self.T1} {self.timeUnit}")
        print(f"The T2 time is: {self.T2} {self.timeUnit}")
        print(f"The depolarization error rate is: {self.dpc_errorRate}")
        print(f"The status of the chip is: {self.status}")
