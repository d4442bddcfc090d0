"""ContextCache stand-ins (the boundary of /root/reference/openmmtools/cache.py:214-461 on this path).

The reference keeps an LRU cache of OpenMM Contexts and selects the Platform through it
(``sampler.energy_context_cache = ContextCache(platform=...)``, multistatesampler.py:1755-1764).  The B200
engine keeps all replicas resident on the device, so there is nothing to cache; these objects only carry the
device selection (``platform`` may be ``None``, ``'CUDA'`` or an object with ``getName()``; ``platform_properties``
may hold ``{'DeviceIndex': '0'}``) so user code that passes them keeps working.
"""


class ContextCache:
    def __init__(self, platform=None, platform_properties=None, **kwargs):
        self._platform = platform
        self._platform_properties = platform_properties
        self.capacity = kwargs.get('capacity', None)
        self.time_to_live = kwargs.get('time_to_live', None)

    @property
    def platform(self):
        return self._platform

    @platform.setter
    def platform(self, p):
        self._platform = p

    @property
    def device_index(self):
        props = self._platform_properties or {}
        for key in ('DeviceIndex', 'CudaDeviceIndex'):
            if key in props:
                return int(str(props[key]).split(',')[0])
        return None

    def get_context(self, thermodynamic_state, integrator=None):
        raise NotImplementedError('there are no OpenMM Contexts on the B200 path; replicas are resident on the GPU')

    def empty(self):
        pass

    def __len__(self):
        return 0


class DummyContextCache(ContextCache):
    pass


global_context_cache = ContextCache()
