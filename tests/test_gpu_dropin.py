"""Drop-in acceptance on the GPU: the reference's own self-test program (main.c), built in the
build container against include/micro_aes.h and linked against libmicro_aes_<bits>.so in place of
micro_aes.c (oracle/Makefile target `dropin`), must print PASSED for every hot-path mode."""
import os
import subprocess

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits,modes", [(128, ["ECB", "CTR", "XTS", "GCM"]), (256, ["XTS", "GCM"])])
def test_reference_main_c_passes_against_our_library(bits, modes):
    exe = os.path.join(ROOT, "oracle", "_ref", f"dropin_main_{bits}")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_main_* not built (needs /root/reference at build time)")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
    assert "FAILED" not in out, out
    for m in modes:
        assert f"AES-{bits} {m} encryption: PASSED!" in out, out
        assert f"AES-{bits} {m} decryption: PASSED!" in out, out
    if bits == 128:      # the header enables OCB and GCM_SIV too: main.c:204-224 and the RFC extras :262-299
        assert out.count("OCB encryption: PASSED!") == 2 and out.count("OCB decryption: PASSED!") == 2, out
        assert out.count("GCMSIV encrypt: PASSED!") == 3 and out.count("GCMSIV decrypt: PASSED!") == 3, out
        assert "CCM encryption: PASSED!" in out and "CCM decryption: PASSED!" in out, out     # main.c:198-204
        assert "EAX encryption: PASSED!" in out and "EAX decryption: PASSED!" in out, out     # main.c:225-237
        assert out.count("SIV encryption: PASSED!") == 3 and out.count("SIV decryption: PASSED!") == 3, out   # :212-218, 300-321
