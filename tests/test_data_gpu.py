"""pai_b200.data.preprocess (dataset.py:51-61,126-134 on the device) against the reference's own host transform:
torchvision ``Resize((256, 256), antialias=True)`` -> ``ConvertImageDtype(float32)`` -> ``Normalize(0.5, 0.5)`` of decoded
grayscale uint8 images.  Bound: one uint8 level (2/255 after normalisation) on every pixel -- torchvision's uint8 resize
uses fixed-point weights on the CPU, so its rounding can differ by one level from the fp32 filter -- and 0.1 level on average."""
import pytest
import torch
from torchvision import transforms

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w", [(256, 256), (512, 512), (300, 417), (1024, 768), (100, 130), (255, 257)])
def test_preprocess_matches_the_reference_transform(h, w):
    from pai_b200 import data
    g = torch.Generator().manual_seed(h * 7 + w)
    yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    smooth = 127 + 90 * torch.sin(yy / 17.0) * torch.cos(xx / 23.0)
    img = (smooth + 25 * torch.randn(h, w, generator=g)).clamp(0, 255).to(torch.uint8)[None]      # [1, H, W] like read_image(GRAY)
    ref_t = transforms.Compose([transforms.Resize((256, 256), antialias=True), transforms.ConvertImageDtype(torch.float32),
                                transforms.Normalize((0.5,), (0.5,))])
    want = ref_t(img)                                     # [1, 256, 256]
    got = data.preprocess(img[None].cuda())[0].cpu()      # [1, 256, 256]
    assert got.shape == want.shape == (1, 256, 256)
    d = (got - want).abs()
    assert float(d.max()) <= 2.0 / 255 + 1e-6, float(d.max()) * 255 / 2
    assert float(d.mean()) <= 0.1 * 2.0 / 255
    if (h, w) == (256, 256):                              # identity resize: exact
        assert torch.equal(got, want)


def test_preprocess_pairs_builds_a_training_batch():
    from pai_b200 import data
    g = torch.Generator().manual_seed(0)
    ins = [torch.randint(0, 256, (1, 300, 300), dtype=torch.uint8, generator=g), torch.randint(0, 256, (1, 256, 512), dtype=torch.uint8, generator=g)]
    gts = [torch.randint(0, 256, (1, 300, 300), dtype=torch.uint8, generator=g), torch.randint(0, 256, (1, 256, 512), dtype=torch.uint8, generator=g)]
    x, t = data.preprocess_pairs(ins, gts)
    assert x.shape == t.shape == (2, 1, 256, 256) and x.is_cuda and x.dtype == torch.float32
    assert float(x.min()) >= -1 and float(x.max()) <= 1
