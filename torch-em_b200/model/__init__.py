from .unet import AnisotropicUNet, UNet2d, UNet3d

__all__ = ["AnisotropicUNet", "UNet2d", "UNet3d"]
