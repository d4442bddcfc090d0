"""openmmtools_b200 -- a B200-native replica-exchange engine behind the openmmtools multistate API.

Only the hot path of choderalab/openmmtools' ``multistate.ReplicaExchangeSampler`` is provided
(mix -> propagate -> energies); module and class names mirror the reference so user
code for that path ports by changing the import.
"""
__version__ = '0.1.0'
