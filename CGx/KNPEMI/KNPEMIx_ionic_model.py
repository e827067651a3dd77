"""Reference import path src/CGx/KNPEMI/KNPEMIx_ionic_model.py -> B200-native membrane model selectors."""
from cgx_b200.ionic_models import (IonicModel, PassiveModel, KirNaKPumpModel, GlialCotransporters,  # noqa: F401
                                   NeuronalCotransporters, ATPPump, HodgkinHuxley)
