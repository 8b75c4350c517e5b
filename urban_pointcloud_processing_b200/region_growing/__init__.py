from .label_connected_comp import LabelConnectedComp
from .layer_lcc import LayerLCC

__all__ = ['LabelConnectedComp', 'LayerLCC']
