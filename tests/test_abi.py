"""C-ABI checks that need no GPU: the shared library loads and exports every symbol include/pcrl.h declares."""
import ctypes
import os

from pointcloud_rl_b200._lib import HEADER, LIB_PATH, lib, parse_header


def test_header_declares_expected_entry_points():
    protos = parse_header(HEADER)
    for name in ("pcrl_stage_points", "pcrl_pointnet_fwd_f32", "pcrl_pointnet_fwd_bf16", "pcrl_pointnet_pack_weights",
                 "pcrl_pointnet_bwd", "pcrl_linear_fwd", "pcrl_linear_bwd", "pcrl_layernorm_fwd", "pcrl_layernorm_bwd",
                 "pcrl_tanh_gaussian_fwd", "pcrl_tanh_gaussian_bwd", "pcrl_td_target", "pcrl_critic_loss",
                 "pcrl_actor_loss", "pcrl_adam_step", "pcrl_polyak", "pcrl_create", "pcrl_destroy", "pcrl_tf32_fallbacks",
                 "pcrl_set_strict_tf32", "pcrl_pointnet_fwd_tf32", "pcrl_pointnet_pack_weights_part",
                 "pcrl_color_jitter_points"):
        assert name in protos, name
    # plain-C signatures only: pointers, fixed-width ints, floats
    for name, (_, args) in protos.items():
        for ctype, _ in args:
            assert ctype in (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32,
                             ctypes.c_int32, ctypes.c_float), (name, ctype)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB_PATH), "build with `python -m pointcloud_rl_b200.build`"
    L = lib()  # binds every prototype; AttributeError if a symbol is missing
    assert L.cdll.pcrl_abi_version() == 1
    assert isinstance(L.last_error(), str)
    # pure-host queries work without a GPU
    img = 128 * 32 + 128 * 128 * 2 + 256 * 128 * 2 + (256 + 512) * 4  # W0' | W1 | signed/permuted W2 | LN parameters
    gen1 = img + 256 * 4 + 16 + 256 * 128 * 2  # + permutation, n_pos, plain W2 (the backward's recompute image)
    gen1 = (gen1 + 127) // 128 * 128
    # second-generation forward image: W0' | centred W1 | Gram(W2c) | centred signed W2 | g1 be1 | g2 be2 | column means
    gen2 = 128 * 32 + 128 * 128 * 2 + 128 * 128 * 2 + 256 * 128 * 2 + 2 * 128 * 4 + 2 * 256 * 4 + (128 + 128) * 4
    gen2 = (gen2 + 127) // 128 * 128
    assert L.pointnet_wpack_bytes(128, 128, 256) == gen1 + gen2
    assert L.pointnet_fwd_f32_workspace(2, 1280, 128, 128, 256) == 2 * 1280 * 512 * 4
    assert L.pointnet_bwd_workspace(4, 256, 128, 128, 256, 8) > 0


def test_sass_uses_blackwell_tensor_and_tma_paths():
    """The shipped cubin must contain tcgen05 MMA (UTC*MMA), TMEM loads (LDTM) and bulk-TMA copies (UBLKCP)."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest

        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UBLKCP" in sass
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", LIB_PATH], capture_output=True, text=True).stdout


def test_host_side_helpers_run_without_a_gpu():
    """pcrl_host_memcpy_mt is pure host code (the staging half of the batch upload): exact bytes for every size / thread
    count, including the single-threaded small-copy path; the size queries of the peer-memory all-reduce are host-only."""
    import numpy as np

    L = lib()
    rng = np.random.default_rng(0)
    for nbytes in (0, 1, 63, 4096, (256 << 10) - 1, 256 << 10, (1 << 20) + 13, 5_000_003):
        src = rng.integers(0, 256, size=nbytes + 64, dtype=np.uint8)
        for threads in (1, 2, 4, 7):
            dst = np.full(nbytes + 64, 0xA5, dtype=np.uint8)
            rc = L.cdll.pcrl_host_memcpy_mt(ctypes.c_void_p(dst.ctypes.data + 3), ctypes.c_void_p(src.ctypes.data + 5),
                                            ctypes.c_int64(nbytes), ctypes.c_int(threads))
            assert rc == 0
            assert np.array_equal(dst[3:3 + nbytes], src[5:5 + nbytes]), (nbytes, threads)
            assert (dst[:3] == 0xA5).all() and (dst[3 + nbytes:] == 0xA5).all(), (nbytes, threads)
    assert L.cdll.pcrl_host_memcpy_mt(None, None, ctypes.c_int64(16), ctypes.c_int(2)) != 0  # NULL with bytes: rejected
    assert L.p2p_flag_bytes() == 16 * 2 * 16 * 4 and L.p2p_state_bytes() == 16 * 4 * 4


def test_host_memcpy_pool_survives_fork():
    """The staging pool's worker threads do not exist in a forked child (the reference forks rollout workers): the child
    must get its own pool instead of waiting for the parent's threads."""
    import numpy as np

    L = lib()
    n = 2 << 20
    src = np.arange(n, dtype=np.uint8)
    dst = np.zeros(n, dtype=np.uint8)
    assert L.cdll.pcrl_host_memcpy_mt(ctypes.c_void_p(dst.ctypes.data), ctypes.c_void_p(src.ctypes.data), ctypes.c_int64(n),
                                      ctypes.c_int(4)) == 0  # the parent's pool exists now
    pid = os.fork()
    if pid == 0:
        code = 1
        try:
            import signal

            signal.alarm(20)  # a child that waits for threads it does not have is killed, not left hanging
            d2 = np.zeros(n, dtype=np.uint8)
            rc = L.cdll.pcrl_host_memcpy_mt(ctypes.c_void_p(d2.ctypes.data), ctypes.c_void_p(src.ctypes.data),
                                            ctypes.c_int64(n), ctypes.c_int(4))
            code = 0 if rc == 0 and np.array_equal(d2, src) else 2
        finally:
            os._exit(code)
    _, status = os.waitpid(pid, 0)
    assert os.WIFEXITED(status) and os.WEXITSTATUS(status) == 0, status
    assert np.array_equal(dst, src)
