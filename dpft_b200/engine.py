"""Fused inference pipeline (placeholder until the fused kernels land)."""


class FusedEngine:
    @staticmethod
    def try_create(model):
        return None
