"""Drop this file next to the reference scripts as ``resnet.py`` (1_HistoPathology/ and
5_JointFusion/): ``from resnet import resnet50`` then resolves to the B200 implementation."""
from multimodalbrainsurvival_b200.resnet import *  # noqa: F401,F403
from multimodalbrainsurvival_b200.resnet import (BasicBlock, Bottleneck, ResNet, ResNetProject, RNfour, RNone,  # noqa: F401
                                                 conv3x3, model_urls, resnet50_1channel, resnet50_4channel)
