"""Host-side construction of QCGD states (the host utilities of src/rules/qcgd.hpp that build
inputs: utils::make_graph :214-230, graphs::randomize :114-120, flags::read_state :1119-1137).

Object layout (qcgd.hpp:63-112), n nodes:
    u16 n | u8 left[n] | u8 right[n] | u16 name_begin[n+1] | sub_node names[...]
    sub_node = { i16 hmlz_and_element, i16 right_or_type, 4 bytes padding, u64 hash } (16 bytes)
A fresh graph names node i with the single element i: hmlz = -1 if i == 0 else i + 1, type = -1
(element), hash = i  (qcgd.hpp:40-42).
"""
import ctypes
import ctypes.util

import numpy as np


def graph_size(n_node):
    return 4 + 20 * n_node


def fresh_graph(n_node) -> np.ndarray:
    """bytes of make_graph(n_node): no particles, node i named i"""
    n = n_node
    g = np.zeros(graph_size(n), np.uint8)
    g[0:2] = np.frombuffer(np.uint16(n).tobytes(), np.uint8)
    nb = np.arange(n + 1, dtype=np.uint16)
    g[2 + 2 * n:4 + 4 * n] = np.frombuffer(nb.tobytes(), np.uint8)
    atoms = np.zeros(n, dtype=[("hmlz", "<i2"), ("kind", "<i2"), ("pad", "<u4"), ("hash", "<u8")])
    idx = np.arange(n)
    atoms["hmlz"] = np.where(idx == 0, -1, idx + 1)
    atoms["kind"] = -1
    atoms["hash"] = idx
    g[4 + 4 * n:] = np.frombuffer(atoms.tobytes(), np.uint8)
    return g


def random_graphs(n_node, n_graphs, seed=0, density=0.5):
    """n_graphs n_node-node graphs with independent random particles (numpy generator): returns
    (sizes u32[n], data u8[n * size]) in the packed interchange form"""
    rng = np.random.default_rng(seed)
    size = graph_size(n_node)
    data = np.tile(fresh_graph(n_node), (n_graphs, 1))
    data[:, 2:2 + 2 * n_node] = rng.random((n_graphs, 2 * n_node), dtype=np.float32) < density
    return np.full(n_graphs, size, np.uint32), data.reshape(-1)


def random_graphs_glibc(n_node, n_graphs, seed=0):
    """same, drawing the particles exactly like flags::read_state + utils::randomize do:
    srand(seed), then per graph and node left = rand() & 1, right = rand() & 1 (qcgd.hpp:114-120)"""
    libc = ctypes.CDLL(ctypes.util.find_library("c"))
    libc.srand(ctypes.c_uint(seed))
    size = graph_size(n_node)
    data = np.tile(fresh_graph(n_node), (n_graphs, 1))
    bits = np.fromiter((libc.rand() & 1 for _ in range(2 * n_node * n_graphs)), np.uint8, 2 * n_node * n_graphs).reshape(n_graphs, n_node, 2)
    data[:, 2:2 + n_node] = bits[:, :, 0]
    data[:, 2 + n_node:2 + 2 * n_node] = bits[:, :, 1]
    return np.full(n_graphs, size, np.uint32), data.reshape(-1)


def read_state_magnitude(n_graphs, real=1.0, imag=0.0):
    """the float arithmetic of flags::read_state (qcgd.hpp:1126-1127)"""
    s = np.sqrt(np.float32(n_graphs))
    return float(np.float32(real) / s), float(np.float32(imag) / s)
