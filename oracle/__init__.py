"""TEST INFRASTRUCTURE ONLY — CPU oracle for the SAMAudio.separate() hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker / the CPU reference being
timed.  The product package (``sam_audio_b200``) never imports it.

Contents
--------
``restate.py``     fp32 torch restatement of the reference algorithm (each
                   function cites the reference file:line it follows).
``ref_loader.py``  imports the *unmodified* reference from /root/reference with
                   ``sys.modules`` stubs for its absent third-party deps.  Only
                   usable in the build container (the GPU box has no
                   /root/reference); used by ``make_golden.py`` to pin
                   ``restate.py`` and to generate ``tests/golden/*.pt``.
``make_golden.py`` the committed generator of the golden fixtures.

Parity status
-------------
* DiT / SAMAudio.forward / align / anchors / processor / ODE control flow:
  PINNED — restate.py is checked against the reference's own modules run in
  this container (see tests/test_oracle_golden.py and make_golden.py).
* DAC-VAE codec arithmetic (third-party ``dacvae @ main``, source absent from
  /root/reference): PARITY UNPINNED by the reference.  The encoder / decoder
  trunks are restated from the published Descript-DAC layout and the reference's
  call sites (codec.py:45-89, config.py:10-41) and are pinned against an
  independent implementation of that layout, ``transformers.models.dac``
  (DacEncoder / DacDecoder, same weights, float64 agreement 1e-12:
  tests/test_oracle_codec_vs_hf_dac.py).  Still unpinned: dacvae's bottleneck
  convention beyond the reference's own call site (in_proj -> first half = mean,
  codec.py:68) and anything dacvae adds on top of the Descript layout.
* T5 numerics: the oracle IS ``transformers.T5EncoderModel`` (installed).
  torchdiffeq (absent): fixed-grid midpoint restated, 32 evaluations at exact
  multiples of 1/32 (model.py:22,285-290); euler and rk4 (3/8 rule) restated
  from the published fixed-grid formulas — PARITY UNPINNED by the reference.
* Visual prompting front end (vision_encoder.py:47-113): the frame transform is
  restated from ATen's antialiased bicubic and PINNED twice — against the
  reference's own PerceptionEncoder class (golden vision.pt, third-party CLIP
  tower replaced by a stand-in) and against torchvision directly
  (tests/test_oracle_vision.py: 2-4 uint8 levels in 1e5 differ by one, the rest
  bit-equal).  The PE-Core tower itself is third-party: unpinned, not built.
* Candidate selection with attached rankers (model.py:306-330): PINNED by a
  golden of the reference's separate() with a fixed-score stand-in ranker.
"""
