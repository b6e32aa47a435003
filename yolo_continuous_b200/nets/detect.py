"""Plain Detect head (reference nets/detect.py): three 1x1 convs, raw NCHW maps, P5 first.

This is the head the shipped YAMLs use (Variant A).  Its convolutions stay
torch modules (cuDNN) -- the B200 kernels take over from `detect.decode_box` onwards.
"""
from torch import nn


class Detect(nn.Module):
    def __init__(self, num_classes=80, anchors=(), ch=()):
        super().__init__()
        self.num_classes = num_classes
        self.len_output = num_classes + 5
        self.num_layers = len(anchors)
        self.num_anchors_each_layer = len(anchors[0]) // 2
        n = self.num_anchors_each_layer * self.len_output
        self.yolo_head_P3 = nn.Conv2d(ch[0], n, 1)
        self.yolo_head_P4 = nn.Conv2d(ch[1], n, 1)
        self.yolo_head_P5 = nn.Conv2d(ch[2], n, 1)
        for mod in self.modules():  # reference nets/detect.py:18-25
            if isinstance(mod, (nn.Conv2d, nn.Linear)):
                nn.init.normal_(mod.weight, 0, 0.01)

    def forward(self, x):
        return [self.yolo_head_P5(x[2]), self.yolo_head_P4(x[1]), self.yolo_head_P3(x[0])]
