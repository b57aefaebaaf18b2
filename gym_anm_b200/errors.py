"""Exception types of the batched ANM environment.

Same names and hierarchy as the reference so that `except` clauses written against
gym-anm keep working (reference gym_anm/errors.py:1-52 and
gym_anm/simulator/components/errors.py:1-62).
"""


# ---- network-dict errors (raised while compiling a network) -----------------------
class InputNetworkFileError(Exception):
    """The network dictionary is malformed."""

    def __init__(self, message=""):
        super().__init__(message)


class BaseMVAError(InputNetworkFileError):
    def __init__(self):
        super().__init__("The network baseMVA should be > 0.")


class BranchSpecError(InputNetworkFileError):
    pass


class BusSpecError(InputNetworkFileError):
    pass


class DeviceSpecError(InputNetworkFileError):
    pass


class GenSpecError(DeviceSpecError):
    pass


class LoadSpecError(DeviceSpecError):
    pass


class StorageSpecError(DeviceSpecError):
    pass


class PFEError(Exception):
    pass


class UnitConversionError(Exception):
    def __init__(self, old, new):
        super().__init__("Cannot convert from %s units to %s units" % (old, new))


# ---- environment-construction errors ------------------------------------------------
class ANMEnvConfigurationError(Exception):
    pass


class ArgsError(ANMEnvConfigurationError):
    pass


class ObsSpaceError(ANMEnvConfigurationError):
    pass


class ObsNotSupportedError(ObsSpaceError):
    def __init__(self, wanted, allowed):
        super().__init__("Observation type unsupported. Desired {} but we only support {}.".format(wanted, allowed))


class UnitsNotSupportedError(ObsSpaceError):
    def __init__(self, wanted, allowed, key):
        super().__init__(
            "Observation unit unsupported. Desired: {} but we only support {} for observation {}.".format(
                wanted, allowed, key
            )
        )


class EnvInitializationError(ANMEnvConfigurationError):
    pass


class EnvNextVarsError(ANMEnvConfigurationError):
    pass


# ---- native-library errors (new) ----------------------------------------------------
class NativeLibraryError(RuntimeError):
    """The CUDA extension is missing or a C-ABI call failed.  There is NO CPU fallback."""


class LPSolverError(RuntimeError):
    """The batched LP solver (include/anm_lp.h) left instances without a usable solution (after a second solve from
    scratch).  Raised by `MPCAgent.act_device` unless the agent was built with `on_lp_failure="host"`."""
