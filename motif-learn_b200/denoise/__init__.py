"""Drop-in subset of ``mtflearn.denoise``: the patch-SVD denoiser ("next" row f4 of SURVEY.md section 8)."""
from ._denoise_svd import DenoiseSVD, denoise_svd, extract_patches, reconstruct_patches

__all__ = ["DenoiseSVD", "denoise_svd", "extract_patches", "reconstruct_patches"]
