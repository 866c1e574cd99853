from .unet import AnisotropicUNet, UNet3d

__all__ = ["AnisotropicUNet", "UNet3d"]
