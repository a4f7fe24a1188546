from typing import Optional, Tuple, Union
from torch import Tensor
PairTensor = Tuple[Tensor, Tensor]
OptPairTensor = Tuple[Tensor, Optional[Tensor]]
OptTensor = Optional[Tensor]
Adj = Tensor
Size = Optional[Tuple[int, int]]
