"""multistate samplers on the B200 engine (mirrors openmmtools.multistate for the replica-exchange path)."""
from .multistatesampler import MultiStateSampler
from .replicaexchange import ReplicaExchangeSampler
from .paralleltempering import ParallelTemperingSampler
from .sams import SAMSSampler
from .utils import SimulationNaNError
from .multistatereporter import MultiStateReporter
