"""owl.elewise -- element-wise operations (reference: owl/owl/elewise.py:6-87)."""
from .narray import NArray


def mult(x, y):
    return NArray.mult(x, y)


def exp(x):
    return NArray.exp(x)


def ln(x):
    return NArray.ln(x)


def sigm(x):
    return NArray.sigm(x)


def relu(x):
    return NArray.relu(x)


def tanh(x):
    return NArray.tanh(x)


def sigm_back(y):
    return NArray.sigm_back(y, y, y)


def relu_back(y, x):
    return NArray.relu_back(y, x, x)      # top aliases bottom, as in the reference (elewise.py:70-78)


def tanh_back(y):
    return NArray.tanh_back(y, y, y)
