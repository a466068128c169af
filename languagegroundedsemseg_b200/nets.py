"""Res16UNet topologies of the hot path, table-driven, over any engine exposing the MinkowskiEngine layer names.

These exist because /root/reference is not present on the GPU box: bench.py / smoke() / -m gpu tests need the
Res16UNet34C graph without importing the reference.  The graphs, parameter names (state-dict ABI) and construction
order (hence RNG consumption at init) match the reference classes
    Res16UNet14A / 34C            models/res16unet.py:298-299, 346-347  (topology :25-270)
    Res16UNet34CR / 34CR_Proj / 34D   models/clip_models.py:95-215
    BasicBlock / NoReluBlock      models/modules/resnet_block.py:10-57, 134-161
    _make_layer                   models/resnet.py:84-125
which tests/test_reference_models_on_facade.py checks key-for-key and value-for-value, tests/test_host_logic.py
(test_reference_models_run_through_the_facade_like_nets) call-for-call through the C ABI, and tests/golden/make_golden.py
output-for-output through the oracle (all in the build container, where the reference is mounted).  With `engine=None` the CUDA facade is used.
"""
import torch
import torch.nn as nn

# (PLANES, LAYERS, flavour) — flavour: 'seg' returns (logits ST, feature ST); 'repr' = CLIP pre-training variants
SPECS = {
    "Res16UNet14A": ((32, 64, 128, 256, 128, 128, 96, 96), (1, 1, 1, 1, 1, 1, 1, 1), "seg"),
    "Res16UNet18A": ((32, 64, 128, 256, 128, 128, 96, 96), (2, 2, 2, 2, 2, 2, 2, 2), "seg"),
    "Res16UNet34A": ((32, 64, 128, 256, 256, 128, 64, 64), (2, 3, 4, 6, 2, 2, 2, 2), "seg"),
    "Res16UNet34C": ((32, 64, 128, 256, 256, 128, 96, 96), (2, 3, 4, 6, 2, 2, 2, 2), "seg"),
    "Res16UNet34CR": ((32, 64, 128, 256, 256, 128, 96, 96), (2, 3, 4, 6, 2, 2, 2, 2), "repr"),
    "Res16UNet34CR_Proj": ((32, 64, 128, 256, 256, 128, 96, 96), (2, 3, 4, 6, 2, 2, 2, 2), "proj"),
    "Res16UNet34D": ((32, 64, 128, 256, 256, 256, 256, 512), (2, 3, 4, 6, 2, 2, 2, 2), "repr"),
}
INIT_DIM = 32


def _engine(engine):
    if engine is None:
        from . import minkowski as engine
    return engine


class ResidualBlock(nn.Module):
    """conv3-BN-ReLU-conv3-BN (+ 1x1 projection of the skip) + add (+ ReLU unless `final_relu` is off)."""

    def __init__(self, E, inplanes, planes, downsample=None, final_relu=True):
        super().__init__()
        self.conv1 = E.MinkowskiConvolution(inplanes, planes, kernel_size=3, stride=1, dilation=1, dimension=3)
        self.norm1 = E.MinkowskiBatchNorm(planes, momentum=0.1)   # _make_layer does not forward bn_momentum
        self.conv2 = E.MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dilation=1, dimension=3)
        self.norm2 = E.MinkowskiBatchNorm(planes, momentum=0.1)
        self.relu = E.MinkowskiReLU(inplace=True)
        self.downsample = downsample
        self.final_relu = final_relu

    def forward(self, x):
        out = self.relu(self.norm1(self.conv1(x)))
        out = self.norm2(self.conv2(out))
        out += x if self.downsample is None else self.downsample(x)
        return self.relu(out) if self.final_relu else out


class Res16UNet(nn.Module):
    def __init__(self, name, in_channels, out_channels, config, D=3, engine=None):
        super().__init__()
        if D != 3:
            raise NotImplementedError("hot path is D = 3")
        E = self.E = _engine(engine)
        planes, layers, flavour = SPECS[name]
        self.name, self.flavour, self.D = name, flavour, D
        self.in_channels, self.out_channels, self.planes = in_channels, out_channels, planes
        bn_m = config.bn_momentum
        c1 = getattr(config, "conv1_kernel_size", 3)

        def conv(i, o, ks, s=1, bias=False):
            return E.MinkowskiConvolution(i, o, kernel_size=ks, stride=s, dilation=1, bias=bias, dimension=3)

        def conv_tr(i, o):
            return E.MinkowskiConvolutionTranspose(i, o, kernel_size=2, stride=2, dilation=1, bias=False, dimension=3)

        def norm(c, m=bn_m):
            return E.MinkowskiBatchNorm(c, momentum=m)

        state = {"inplanes": INIT_DIM}

        def stage(p, n_blocks):
            blocks = []
            for b in range(n_blocks):
                ds = None
                if b == 0 and state["inplanes"] != p:
                    # the projection branch is created with the stage's bn_momentum (models/resnet.py:94-103)
                    ds = nn.Sequential(conv(state["inplanes"], p, 1), norm(p))
                blocks.append(ResidualBlock(E, state["inplanes"], p, ds))
                state["inplanes"] = p
            return nn.Sequential(*blocks)

        self.conv0p1s1, self.bn0 = conv(in_channels, INIT_DIM, c1), norm(INIT_DIM)
        enc_names = [("conv1p1s2", "bn1", "block1"), ("conv2p2s2", "bn2", "block2"),
                     ("conv3p4s2", "bn3", "block3"), ("conv4p8s2", "bn4", "block4")]
        for lvl, (cn, bn, bl) in enumerate(enc_names):
            setattr(self, cn, conv(state["inplanes"], state["inplanes"], 2, 2))
            setattr(self, bn, norm(state["inplanes"]))
            setattr(self, bl, stage(planes[lvl], layers[lvl]))
        dec_names = [("convtr4p16s2", "bntr4", "block5"), ("convtr5p8s2", "bntr5", "block6"),
                     ("convtr6p4s2", "bntr6", "block7"), ("convtr7p2s2", "bntr7", "block8")]
        skips = [planes[2], planes[1], planes[0], INIT_DIM]
        for lvl, (cn, bn, bl) in enumerate(dec_names):
            p = planes[4 + lvl]
            setattr(self, cn, conv_tr(state["inplanes"], p))
            setattr(self, bn, norm(p))
            state["inplanes"] = p + skips[lvl]
            setattr(self, bl, stage(p, layers[4 + lvl]))
        self.final = conv(planes[7], out_channels, 1, bias=True)
        self.relu = E.MinkowskiReLU(inplace=True)
        if flavour in ("repr", "proj"):
            self.block8[-1].final_relu = False      # NoReluBlock (models/clip_models.py:101)
            self.repr_only = False
            # The reference builds NoReluBlock by constructing a fresh block and then re-pointing its layers at
            # the source block's (resnet_block.py:136-144); 34D does it twice (clip_models.py:208-211).  Draw the
            # same random numbers so that a given torch seed yields identical parameters.
            for _ in range(2 if name == "Res16UNet34D" else 1):
                ResidualBlock(E, planes[7], planes[7])
        if flavour == "proj":
            self.projection_layer = nn.Conv1d(512, planes[7], kernel_size=1, stride=1, bias=True)
        self._enc, self._dec = enc_names, dec_names
        for m in self.modules():                    # models/resnet.py:78-82
            if isinstance(m, E.MinkowskiBatchNorm):
                nn.init.constant_(m.bn.weight, 1)
                nn.init.constant_(m.bn.bias, 0)

    def representation_only(self, flag):            # models/clip_models.py:106-109
        self.repr_only = flag
        self.final = None

    def features(self, x):
        E = self.E
        out = self.relu(self.bn0(self.conv0p1s1(x)))
        skips = [out]
        for lvl, (cn, bn, bl) in enumerate(self._enc):
            out = self.relu(getattr(self, bn)(getattr(self, cn)(out)))
            out = getattr(self, bl)(out)
            if lvl < 3:
                skips.append(out)
        for lvl, (cn, bn, bl) in enumerate(self._dec):
            out = self.relu(getattr(self, bn)(getattr(self, cn)(out)))
            out = E.cat(out, skips[3 - lvl])
            out = getattr(self, bl)(out)
        return out

    def forward(self, x, anchor_feats=None):
        out = self.features(x)
        if self.flavour == "seg":
            return self.final(out), out
        res = out if self.repr_only else (self.final(out), out)
        if self.flavour == "proj":
            return res, self.projection_layer(anchor_feats.unsqueeze(-1)).squeeze()
        return res


def build_model(name, in_channels, out_channels, config, engine=None):
    """`models.load_model(name)(in_channels, out_channels, config)` of the reference (models/__init__.py:33)."""
    if name not in SPECS:
        raise KeyError(f"{name}: not a hot-path topology; available: {sorted(SPECS)}")
    return Res16UNet(name, in_channels, out_channels, config, engine=engine)


class DefaultConfig:
    """The flags the topologies read (config/config.py:54,57,89)."""
    bn_momentum = 0.02
    conv1_kernel_size = 3
    dilations = [1, 1, 1, 1]
