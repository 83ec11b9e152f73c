"""exchanger_b200: B200-native sampler for exchanger's partially-collapsed Gibbs sweep.

Host-side mirror of the reference's R API (``exchanger``, ``run_inference``,
``ExchangERFit``) over the C ABI declared in ``include/exb200.h``.
"""
from .model import (Attribute, AttributeIndex, BetaRV, CategoricalAttribute, DirichletProcess,
                    DirichletRV, EwensRP, ExchangERModel, GammaRV, GeneralizedCouponRP,
                    Levenshtein, PitmanYorRP, ShiftedNegBinomRV, exchanger, transform_dist_fn)

__all__ = [
    "Attribute", "AttributeIndex", "BetaRV", "CategoricalAttribute", "DirichletProcess",
    "DirichletRV", "EwensRP", "ExchangERModel", "GammaRV", "GeneralizedCouponRP",
    "Levenshtein", "PitmanYorRP", "ShiftedNegBinomRV", "exchanger", "transform_dist_fn",
]
