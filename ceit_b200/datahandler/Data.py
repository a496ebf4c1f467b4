"""Recorded measurement sets (host-side mirror of CEIT/datahandler/Data.py).

A recording is an (samples x channels) array; `data_mean` is its mean over the samples and, when the
recording still contains the excitation electrodes' own channels (`contain_excitation`), every channel
i with i % electrode_num == 0 is dropped (Data.py:18-30), which leaves the L(L-1) pattern vector the
inverse model expects.  Pure host logic: nothing here touches the device until `get_capacitance`.
"""
from abc import ABCMeta, abstractmethod

import numpy as np


class BasicData(metaclass=ABCMeta):
    def __init__(self, data, electrode_num=16, contain_excitation=False):
        self.data = data
        self.data_mean = np.mean(self.data, axis=0)
        self.contain_excitation = contain_excitation
        if contain_excitation:
            self.data_mean = self.exclude_excitation(electrode_num)
        self.delta_V = None

    def exclude_excitation(self, electrode_num):
        keep = np.arange(len(self.data_mean)) % electrode_num != 0
        return np.array(self.data_mean[keep])

    @abstractmethod
    def return_copy(self):
        """Returns copy of the current data"""


class CalibData(BasicData):
    def __init__(self, data, i, contain_excitation=False):
        super().__init__(data, contain_excitation=contain_excitation)
        self.height_idx = i

    def return_copy(self):
        return CalibData(np.copy(self.data), self.height_idx, contain_excitation=self.contain_excitation)


class Data(BasicData):
    def __init__(self, data, x, y, z, transformed=False, contain_excitation=False):
        if transformed:
            self.x, self.y, self.z = x, y, z
        else:
            self.x, self.y = translate_corr(x, y)
            self.z = z
        self.capacitance_predict = None
        super().__init__(data, contain_excitation=contain_excitation)

    def calc_delta_V(self, calibration):
        """delta_V = data_mean - calibration mean (Data.py:82-92)."""
        if len(self.data_mean) != len(calibration.data_mean):
            raise Exception("Shape does not match with calibration.")
        self.delta_V = self.data_mean - np.copy(calibration.data_mean)

    def get_capacitance(self, solver):
        """Reconstruct once and cache (Data.py:94-112)."""
        if self.capacitance_predict is not None:
            return np.copy(self.capacitance_predict)
        if self.delta_V is None:
            print("Be aware that you haven't generated delta_V yet, the method would return an empty NDarray")
            return np.array([])
        self.capacitance_predict = solver.solve(self.delta_V)
        return np.copy(self.capacitance_predict)

    def return_copy(self):
        return Data(np.copy(self.data), self.x, self.y, self.z, transformed=True,
                    contain_excitation=self.contain_excitation)


def translate_corr(x, y):
    return x - 100, -y + 100


def rev_translate_corr(x, y):
    return x + 100, - y + 100
