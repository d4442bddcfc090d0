"""Alchemical modification of the Lennard-Jones test system and the AlchemicalState that controls it.

Mirrors /root/reference/openmmtools/alchemy/alchemy.py for the pieces the replica-exchange path exercises:
``AlchemicalState`` (:90-410), ``AlchemicalRegion`` (:417-427) and
``AbsoluteAlchemicalFactory.create_alchemical_system`` (:637-754).  What the factory does to a zero-charge LJ
fluid (:1539-2038) is restated as parameters of the CUDA energy function: non-alchemical pairs keep plain LJ
(NonbondedForce with the alchemical epsilons zeroed, :1903-1909), alchemical/non-alchemical pairs use soft-core
sterics controlled by ``lambda_sterics`` (:1383-1388, interaction group :1915), alchemical/alchemical pairs use
the same expression with lambda fixed to 1 unless ``annihilate_sterics`` (:1776-1777, group :1919).  Bonded,
electrostatic, GB and multi-region branches need a molecular force field and are out of scope.
"""
import collections
from . import unit
from .states import GlobalParameterState, GlobalParameterError   # noqa: F401
from .system import System, LJ


class AlchemicalStateError(GlobalParameterError):
    pass


class AlchemicalState(GlobalParameterState):
    """lambda_sterics, lambda_electrostatics, lambda_bonds, lambda_angles, lambda_torsions in [0, 1]
    (alchemy.py:90-410)."""

    class _LambdaParameter(GlobalParameterState.GlobalParameter):
        def __init__(self, name):
            super().__init__(name, standard_value=1.0, validator=self.lambda_validator)

        @staticmethod
        def lambda_validator(self_, new_value):
            if new_value is None:
                return None
            if not (0.0 <= new_value <= 1.0):      # alchemy.py:213-219
                raise ValueError('{} must be between 0 and 1'.format(new_value))
            return float(new_value)

    lambda_sterics = _LambdaParameter('lambda_sterics')
    lambda_electrostatics = _LambdaParameter('lambda_electrostatics')
    lambda_bonds = _LambdaParameter('lambda_bonds')
    lambda_angles = _LambdaParameter('lambda_angles')
    lambda_torsions = _LambdaParameter('lambda_torsions')

    def set_alchemical_parameters(self, new_value):
        """Set all defined lambda parameters to the given value (alchemy.py:255-269)."""
        for name, v in self._parameters.items():
            if v is not None:
                setattr(self, name, new_value)


_AR_FIELDS = ['alchemical_atoms', 'alchemical_bonds', 'alchemical_angles', 'alchemical_torsions',
              'annihilate_electrostatics', 'annihilate_sterics', 'softcore_alpha', 'softcore_a', 'softcore_b',
              'softcore_c', 'softcore_beta', 'softcore_d', 'softcore_e', 'softcore_f', 'name']
AlchemicalRegion = collections.namedtuple('AlchemicalRegion', _AR_FIELDS)
# defaults as alchemy.py:417-427
AlchemicalRegion.__new__.__defaults__ = (None, None, None, None, True, False, 0.5, 1, 1, 6, 0.0, 1, 1, 2, None)


class AbsoluteAlchemicalFactory:
    """Factory of alchemically modified systems (alchemy.py:430-754), LJ-sterics subset."""

    def __init__(self, consistent_exceptions=False, switch_width=1.0 * unit.angstroms,
                 alchemical_pme_treatment='exact', alchemical_rf_treatment='switched',
                 disable_alchemical_dispersion_correction=False, split_alchemical_forces=True):
        self.consistent_exceptions = consistent_exceptions
        self.switch_width = switch_width
        self.alchemical_pme_treatment = alchemical_pme_treatment
        self.alchemical_rf_treatment = alchemical_rf_treatment
        self.disable_alchemical_dispersion_correction = disable_alchemical_dispersion_correction
        self.split_alchemical_forces = split_alchemical_forces

    def create_alchemical_system(self, reference_system, alchemical_regions, alchemical_regions_interactions=frozenset()):
        if not isinstance(reference_system, System) or reference_system.kind != LJ:
            raise NotImplementedError('only the Lennard-Jones test system can be alchemically modified on this path')
        if isinstance(alchemical_regions, (list, tuple)) and not isinstance(alchemical_regions, AlchemicalRegion):
            if len(alchemical_regions) != 1:
                raise NotImplementedError('multiple alchemical regions are not provided')
            alchemical_regions = alchemical_regions[0]
        region = alchemical_regions
        if region.alchemical_atoms is None or len(list(region.alchemical_atoms)) == 0:
            raise ValueError('alchemical_atoms must be specified')     # alchemy.py:835
        atoms = sorted(set(int(a) for a in region.alchemical_atoms))
        n = reference_system.n_particles
        if atoms[0] < 0 or atoms[-1] >= n:
            raise ValueError('alchemical atom index out of range')
        if any(x for x in (region.alchemical_bonds, region.alchemical_angles, region.alchemical_torsions)
               if x not in (None, False, [], ())):
            raise NotImplementedError('bonded alchemical terms are outside the hot path (the LJ fluid has none)')
        if reference_system.is_alchemical:
            raise ValueError('the system is already alchemically modified')
        system = reference_system.copy()
        system.alchemical_atoms = tuple(atoms)
        system.annihilate_sterics = bool(region.annihilate_sterics)
        system.softcore_alpha = float(region.softcore_alpha)
        system.softcore_a = float(region.softcore_a)
        system.softcore_b = float(region.softcore_b)
        system.softcore_c = float(region.softcore_c)
        # alchemy.py:1787-1790
        system.alchemical_dispersion_correction = (not self.disable_alchemical_dispersion_correction) and \
            reference_system.use_dispersion_correction
        # lambda_electrostatics is defined by the factory even for zero charges (the electrostatic forces exist but
        # are identically zero), so AlchemicalState.from_system finds both parameters
        system.global_parameters = dict(reference_system.global_parameters)
        system.global_parameters['lambda_sterics'] = 1.0
        system.global_parameters['lambda_electrostatics'] = 1.0
        return system

    @staticmethod
    def get_energy_components(alchemical_system, alchemical_state, positions, platform=None):
        raise NotImplementedError('per-force energy dissection needs OpenMM; see tests/ for the component checks')
