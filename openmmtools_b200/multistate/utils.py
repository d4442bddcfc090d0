"""Exceptions of the multistate package (/root/reference/openmmtools/multistate/utils.py:51)."""


class SimulationNaNError(Exception):
    """Error when a simulation goes to NaN."""
    pass
