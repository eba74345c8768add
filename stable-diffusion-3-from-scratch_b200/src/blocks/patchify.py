"""(un)patchify with the reference's function API (reference: src/blocks/patchify.py:4-72).
Column order inside a token is c*p*p + i*p + j, the layout of a flattened conv weight."""
from mmdit.functional import PatchifyFn, UnpatchifyFn


def patchify(images, patch_size):
    ph, pw = patch_size
    N, C, H, W = images.shape
    if ph != pw or H % ph or W % pw:
        raise NotImplementedError("patchify: square patches that divide H and W only (no padding path)")
    return PatchifyFn.apply(images, ph).view(N, (H // ph) * (W // pw), C * ph * pw)


def unpatchify(patches, patch_size, original_shape):
    ph, pw = patch_size
    H, W = original_shape
    N, num_patches, patch_dim = patches.shape
    if ph != pw or H % ph or W % pw:
        raise NotImplementedError("unpatchify: square patches that divide H and W only")
    C = patch_dim // (ph * pw)
    return UnpatchifyFn.apply(patches.reshape(N * num_patches, patch_dim), N, C, H, W, ph)
