"""Import-path shim: exposes the B200 implementation under the reference's module paths
(`src.caldera.utils.dataclasses`, `src.caldera.utils.quantization`,
`src.caldera.decomposition.alg`, `src.convex_caldera.decomposition.convex_caldera`), so code
written against genglongling/EE274_ConvexCaldera_LLM_quantization with
`rank-constrained-regression-main` on sys.path runs unchanged with this repository's root
on sys.path instead."""
