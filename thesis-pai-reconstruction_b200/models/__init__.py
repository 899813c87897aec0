"""Drop-in replacement for the reference's ``models`` package on the Pix2Pix / PatchGAN /
SSIM-PSNR hot path.  Same module paths, class names, constructor signatures,
``state_dict`` keys and loss / metric interfaces as /root/reference/models/{pix2pix,wrapper,utils}.py;
the arithmetic runs on the B200 kernels of ``pai_b200`` (no PyTorch-op fallback for the hot path).

Put this directory's parent ahead of the reference checkout on ``sys.path`` and ``main.py`` /
``report.py`` pick it up unchanged.
"""
