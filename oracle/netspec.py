"""ORACLE (test infrastructure, never on the product path).

Data-driven netlist of the YOLOv4 graph the reference assembles with Keras.
Restates, as a flat op list:
  conv()            /root/reference/custom_layers.py:5-31
  residual_block()  /root/reference/custom_layers.py:34-44   (add AFTER the activation)
  csp_block()       /root/reference/custom_layers.py:47-69   (route conv created FIRST, concat [main, route])
  cspdarknet53()    /root/reference/custom_layers.py:100-138 (SPP concat order [mp13, mp9, mp5, x])
  yolov4_neck()     /root/reference/custom_layers.py:141-198

Conv indices are Keras creation order == darknet file order, which is what
/root/reference/utils.py:12-53 relies on (conv_layer_size = 110, heads 93/101/109).
Parity unpinned: the reference's own runtime (TensorFlow) is absent here, see DESIGN.md.
"""
from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class Op:
    kind: str                      # 'conv' | 'add' | 'concat' | 'maxpool' | 'upsample'
    out: str
    ins: List[str]
    # conv only
    idx: int = -1
    cin: int = 0
    cout: int = 0
    k: int = 0
    stride: int = 1
    bn: bool = True
    act: str = 'linear'            # 'mish' | 'leaky' | 'linear'
    pool: int = 0                  # maxpool only
    channels: int = 0              # output channels (all ops)
    scale: int = 1                 # output spatial size = img_size // scale


class _Builder:
    def __init__(self):
        self.ops: List[Op] = []
        self.nconv = 0
        self.meta = {'img': (3, 1)}   # name -> (channels, scale)
        self.nadd = 0
        self.ncat = 0

    def conv(self, x, filters, k, down=False, act='leaky', bn=True):
        cin, sc = self.meta[x]
        name = f'c{self.nconv}'
        self.ops.append(Op('conv', name, [x], idx=self.nconv, cin=cin, cout=filters, k=k,
                           stride=2 if down else 1, bn=bn, act=act,
                           channels=filters, scale=sc * (2 if down else 1)))
        self.meta[name] = (filters, sc * (2 if down else 1))
        self.nconv += 1
        return name

    def add(self, a, b):
        self.nadd += 1
        name = f'r{self.nadd}'
        ch, sc = self.meta[a]
        assert self.meta[b] == (ch, sc)
        self.ops.append(Op('add', name, [a, b], channels=ch, scale=sc))
        self.meta[name] = (ch, sc)
        return name

    def concat(self, parts, name=None):
        self.ncat += 1
        name = name or f'cat{self.ncat}'
        sc = self.meta[parts[0]][1]
        assert all(self.meta[p][1] == sc for p in parts)
        ch = sum(self.meta[p][0] for p in parts)
        self.ops.append(Op('concat', name, list(parts), channels=ch, scale=sc))
        self.meta[name] = (ch, sc)
        return name

    def maxpool(self, x, size):
        name = f'mp{size}'
        ch, sc = self.meta[x]
        self.ops.append(Op('maxpool', name, [x], pool=size, channels=ch, scale=sc))
        self.meta[name] = (ch, sc)
        return name

    def upsample(self, x):
        name = f'up_{x}'
        ch, sc = self.meta[x]
        assert sc % 2 == 0
        self.ops.append(Op('upsample', name, [x], channels=ch, scale=sc // 2))
        self.meta[name] = (ch, sc // 2)
        return name

    # custom_layers.py:34-44
    def residual(self, x, f1, f2, act):
        y = self.conv(x, f1, 1, act=act)
        y = self.conv(y, f2, 3, act=act)
        return self.add(x, y)

    # custom_layers.py:47-69
    def csp(self, x, out, repeat, bottleneck=False):
        route = self.conv(x, out, 1, act='mish')
        x = self.conv(x, out, 1, act='mish')
        for _ in range(repeat):
            x = self.residual(x, out // 2 if bottleneck else out, out, 'mish')
        x = self.conv(x, out, 1, act='mish')
        return self.concat([x, route])


def build_netlist(num_classes: int = 80):
    """Returns (ops, head_names).  Execution order == list order."""
    b = _Builder()
    # --- cspdarknet53, custom_layers.py:100-138
    x = b.conv('img', 32, 3)                       # leaky (default arg), NOT mish
    x = b.conv(x, 64, 3, down=True)                # leaky
    x = b.csp(x, 64, 1, bottleneck=True)
    x = b.conv(x, 64, 1, act='mish')
    x = b.conv(x, 128, 3, down=True, act='mish')
    x = b.csp(x, 64, 2)
    x = b.conv(x, 128, 1, act='mish')
    x = b.conv(x, 256, 3, down=True, act='mish')
    x = b.csp(x, 128, 8)
    x = b.conv(x, 256, 1, act='mish')
    route0 = x
    x = b.conv(x, 512, 3, down=True, act='mish')
    x = b.csp(x, 256, 8)
    x = b.conv(x, 512, 1, act='mish')
    route1 = x
    x = b.conv(x, 1024, 3, down=True, act='mish')
    x = b.csp(x, 512, 4)
    x = b.conv(x, 1024, 1, act='mish')
    x = b.conv(x, 512, 1)
    x = b.conv(x, 1024, 3)
    x = b.conv(x, 512, 1)
    x = b.concat([b.maxpool(x, 13), b.maxpool(x, 9), b.maxpool(x, 5), x])
    x = b.conv(x, 512, 1)
    x = b.conv(x, 1024, 3)
    route2 = b.conv(x, 512, 1)
    # --- yolov4_neck, custom_layers.py:141-198
    nout = 3 * (num_classes + 5)
    x = b.conv(route2, 256, 1)
    up = b.upsample(x)
    r1 = b.conv(route1, 256, 1)
    x = b.concat([r1, up])
    for f, k in ((256, 1), (512, 3), (256, 1), (512, 3), (256, 1)):
        x = b.conv(x, f, k)
    route1b = x
    x = b.conv(x, 128, 1)
    up = b.upsample(x)
    r0 = b.conv(route0, 128, 1)
    x = b.concat([r0, up])
    for f, k in ((128, 1), (256, 3), (128, 1), (256, 3), (128, 1)):
        x = b.conv(x, f, k)
    route0b = x
    x = b.conv(x, 256, 3)
    head_s = b.conv(x, nout, 1, act='linear', bn=False)
    x = b.conv(route0b, 256, 3, down=True)
    x = b.concat([x, route1b])
    for f, k in ((256, 1), (512, 3), (256, 1), (512, 3), (256, 1)):
        x = b.conv(x, f, k)
    route1c = x
    x = b.conv(x, 512, 3)
    head_m = b.conv(x, nout, 1, act='linear', bn=False)
    x = b.conv(route1c, 512, 3, down=True)
    x = b.concat([x, route2])
    for f, k in ((512, 1), (1024, 3), (512, 1), (1024, 3), (512, 1)):
        x = b.conv(x, f, k)
    x = b.conv(x, 1024, 3)
    head_l = b.conv(x, nout, 1, act='linear', bn=False)
    return b.ops, [head_s, head_m, head_l]


def conv_ops(num_classes: int = 80):
    return [o for o in build_netlist(num_classes)[0] if o.kind == 'conv']


def conv_gflop(img_size: int, num_classes: int = 80) -> float:
    tot = 0
    for o in conv_ops(num_classes):
        hw = img_size // o.scale
        tot += 2 * hw * hw * o.cout * o.cin * o.k * o.k
    return tot / 1e9


def darknet_file_floats(num_classes: int = 80) -> int:
    n = 0
    for o in conv_ops(num_classes):
        n += (4 * o.cout if o.bn else o.cout) + o.cout * o.cin * o.k * o.k
    return n
