"""A small UNet-shaped module tree used by the model-level parity fixture: the names exercise the skip-key rules of the reference
(common_skip_keys: `.time_embed*`, `proj_out`, ...), Linear / Conv2d / ConvTranspose2d / Embedding / LayerNorm leaves, small layers
below the numel / channel thresholds, and nested containers.  Shared by generate_model.py (reference side) and the test."""
import torch


class Block(torch.nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = torch.nn.LayerNorm(c)
        self.to_q = torch.nn.Linear(c, c, bias=False)
        self.to_k = torch.nn.Linear(2 * c, c, bias=False)
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(c, c), torch.nn.Dropout(0.0)])
        self.ff = torch.nn.Sequential(torch.nn.Linear(c, 4 * c), torch.nn.GELU(), torch.nn.Linear(4 * c, c))


class Toy(torch.nn.Module):
    def __init__(self, c=128):
        super().__init__()
        self.time_embedding = torch.nn.Sequential(torch.nn.Linear(32, c), torch.nn.SiLU(), torch.nn.Linear(c, c))
        self.conv_in = torch.nn.Conv2d(4, c, 3, padding=1)
        self.down = torch.nn.ModuleList([torch.nn.Conv2d(c, c, 3, padding=1), torch.nn.Conv2d(c, c, 3, stride=2, padding=1)])
        self.mid = torch.nn.ModuleList([Block(c), Block(c)])
        self.up = torch.nn.ConvTranspose2d(c, c, 4, stride=2, padding=1)
        self.tiny = torch.nn.Linear(16, 16)
        self.token_embedding = torch.nn.Embedding(512, c)
        self.conv_out = torch.nn.Conv2d(c, 4, 3, padding=1)
        self.proj_out = torch.nn.Linear(c, c)


def build(seed=0):
    torch.manual_seed(seed)
    return Toy().to(torch.bfloat16)


CONFIGS = {
    "int8_default": dict(weights_dtype="int8"),
    "int8_w8a8_conv": dict(weights_dtype="int8", use_quantized_matmul=True, quant_conv=True, use_quantized_matmul_conv=True),
    "uint4_conv_embedding": dict(weights_dtype="uint4", quant_conv=True, quant_embedding=True),
    "fp8_hadamard_w8a8": dict(weights_dtype="float8_e4m3fn", use_quantized_matmul=True, use_hadamard=True),
    "int6_no_skip_keys": dict(weights_dtype="int6", add_skip_keys=False, quant_conv=True),
    "int8_explicit_skips": dict(weights_dtype="int8", modules_to_not_convert=["mid.1", "down.0"], modules_dtype_dict={"uint4": ["ff.0"]},
                                use_quantized_matmul=True, modules_to_not_use_matmul=["to_q"]),
}
