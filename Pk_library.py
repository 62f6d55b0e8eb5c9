"""`import Pk_library as PKL` -- drop-in name of the reference package (library/Pk_library/__init__.py:1-4:
`from Pk_library import *; from Pk_snapshot import *; from bispectrum_library import *`)."""
from pylians_b200.Pk_library import *  # noqa: F401,F403
from pylians_b200.Pk_library import (Pk, XPk, frequencies, MAS_function, MAS_correction, check_number_modes, FFT3Dr_f,  # noqa: F401
                                     Pk_plane, XPk_plane, Pk_theta, correct_MAS, Xi, frequencies_2D, check_number_modes_2D,
                                     FFT2Dr_f, IFFT2Dr_f, IFFT3Dr_f, XPk_imag)
from pylians_b200.Pk_snapshot import Pk_comp, Pk_Gadget, name_dict  # noqa: F401
from pylians_b200.bispectrum_library import Bk, F2, Bispectrum_theory  # noqa: F401
