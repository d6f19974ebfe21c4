"""U^2-Net (small, "U2NETP") for saliency inputs, hook-site compatible.

Scaffolding for the scoring path.  One depth-parameterised residual U-block
replaces the reference's five hand-unrolled classes
(/root/reference/models/DUTS/u2net.py:31-383) while keeping every attribute
name the hook-site table addresses (``stageK.rebnconvin.relu_s1``,
``stageK.rebnconvN.relu_s1``, ``stageK.rebnconvNd.relu_s1``, ``sideK``), the
``int((1-rate)*ch)`` clamped-to-1 channel rule (u2net.py:37-52, 437-453) and the
parameter-creation order (so seeded init matches the reference's).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _kept_clamped(ch, rate):
    return max(1, int((1 - rate) * ch))


class REBNCONV(nn.Module):
    def __init__(self, in_ch=3, out_ch=3, dirate=1):
        super().__init__()
        self.conv_s1 = nn.Conv2d(in_ch, out_ch, 3, padding=dirate, dilation=dirate)
        self.bn_s1 = nn.BatchNorm2d(out_ch)
        self.relu_s1 = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.relu_s1(self.bn_s1(self.conv_s1(x)))


def _upsample_like(src, tar):
    return F.interpolate(src, size=tar.shape[2:], mode='bilinear', align_corners=False)


class RSU(nn.Module):
    """Residual U-block of `depth` encoder convs (7, 6, 5, 4) or the dilated
    pool-free variant (`dilated=True`, the reference's RSU4F)."""

    def __init__(self, depth, compress_rate, in_ch, mid_ch, out_ch, dilated=False):
        super().__init__()
        self.depth, self.dilated = depth, dilated
        # widths[i] = output width of encoder conv i+1; last two are never pruned
        widths = [_kept_clamped(mid_ch, compress_rate[i]) for i in range(depth - 2)] + [mid_ch, mid_ch]
        self.rebnconvin = REBNCONV(in_ch, out_ch, 1)
        cin = out_ch
        for i in range(1, depth + 1):
            if dilated:
                rate = 2 ** (i - 1)
            else:
                rate = 2 if i == depth else 1
            setattr(self, 'rebnconv%d' % i, REBNCONV(cin, widths[i - 1], rate))
            cin = widths[i - 1]
            if not dilated and i <= depth - 2:
                setattr(self, 'pool%d' % i, nn.MaxPool2d(2, stride=2, ceil_mode=True))
        for i in range(depth - 1, 0, -1):
            cout = widths[i - 2] if i > 1 else out_ch
            rate = 2 ** (i - 1) if dilated else 1
            setattr(self, 'rebnconv%dd' % i, REBNCONV(widths[i - 1] * 2, cout, rate))

    def forward(self, x):
        hxin = self.rebnconvin(x)
        enc, hx = [], hxin
        for i in range(1, self.depth + 1):
            hx = getattr(self, 'rebnconv%d' % i)(hx)
            enc.append(hx)
            if not self.dilated and i <= self.depth - 2:
                hx = getattr(self, 'pool%d' % i)(hx)
        d = enc[-1]
        for i in range(self.depth - 1, 0, -1):
            skip = enc[i - 1]
            if d.shape[2:] != skip.shape[2:]:
                d = _upsample_like(d, skip)
            d = getattr(self, 'rebnconv%dd' % i)(torch.cat((d, skip), 1))
        return d + hxin


def u2netp_rate_groups(compress_rate):
    """Slice the 39 rates into encoder / decoder / stage-output groups (u2net.py:386-410)."""
    cuts = [0, 5, 9, 12, 14, 16, 18, 23, 27, 30, 32, 34, 39]
    parts = [compress_rate[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    return parts[:6], parts[6:11], parts[11]


class U2NETP(nn.Module):
    def __init__(self, in_ch=3, out_ch=1, compress_rate=[0.] * 100):
        super().__init__()
        self.best_loss = 999999
        enc, dec, exter = u2netp_rate_groups(compress_rate)
        m1, m2, m3, m4, m5 = (_kept_clamped(64, exter[i]) for i in range(5))
        pool = lambda: nn.MaxPool2d(2, stride=2, ceil_mode=True)
        self.stage1 = RSU(7, enc[0], in_ch, 16, m1)
        self.pool12 = pool()
        self.stage2 = RSU(6, enc[1], m1, 16, m2)
        self.pool23 = pool()
        self.stage3 = RSU(5, enc[2], m2, 16, m3)
        self.pool34 = pool()
        self.stage4 = RSU(4, enc[3], m3, 16, m4)
        self.pool45 = pool()
        self.stage5 = RSU(4, enc[4], m4, 16, m5, dilated=True)
        self.pool56 = pool()
        self.stage6 = RSU(4, enc[5], m5, 16, m5, dilated=True)
        self.stage5d = RSU(4, dec[4], m5 * 2, 16, m4, dilated=True)
        self.stage4d = RSU(4, dec[3], m4 * 2, 16, m3)
        self.stage3d = RSU(5, dec[2], m3 * 2, 16, m2)
        self.stage2d = RSU(6, dec[1], m2 * 2, 16, m1)
        self.stage1d = RSU(7, dec[0], m1 * 2, 16, m1)
        for i, width in enumerate((m1, m1, m2, m3, m4, m5)):
            setattr(self, 'side%d' % (i + 1), nn.Conv2d(width, out_ch, 3, padding=1))
        self.outconv = nn.Conv2d(6, out_ch, 1)

    def forward(self, x):
        hx1 = self.stage1(x)
        hx2 = self.stage2(self.pool12(hx1))
        hx3 = self.stage3(self.pool23(hx2))
        hx4 = self.stage4(self.pool34(hx3))
        hx5 = self.stage5(self.pool45(hx4))
        hx6 = self.stage6(self.pool56(hx5))
        hx5d = self.stage5d(torch.cat((_upsample_like(hx6, hx5), hx5), 1))
        hx4d = self.stage4d(torch.cat((_upsample_like(hx5d, hx4), hx4), 1))
        hx3d = self.stage3d(torch.cat((_upsample_like(hx4d, hx3), hx3), 1))
        hx2d = self.stage2d(torch.cat((_upsample_like(hx3d, hx2), hx2), 1))
        hx1d = self.stage1d(torch.cat((_upsample_like(hx2d, hx1), hx1), 1))
        d1 = self.side1(hx1d)
        sides = [d1] + [_upsample_like(getattr(self, 'side%d' % k)(h), d1)
                        for k, h in ((2, hx2d), (3, hx3d), (4, hx4d), (5, hx5d), (6, hx6))]
        d0 = self.outconv(torch.cat(sides, 1))
        return tuple(torch.sigmoid(t) for t in [d0] + sides)


def u2netp(compress_rate=[0.] * 100):
    return U2NETP(compress_rate=compress_rate)
