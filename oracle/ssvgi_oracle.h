/*
 * ssvgi_oracle.h — TEST INFRASTRUCTURE. Entry points of the plain-C CPU restatement of the SSVGI passes
 * (see ssvgi_oracle.c). Signatures are those of include/lgcu.h minus the stream argument; images live in HOST memory.
 */
#ifndef SSVGI_ORACLE_H
#define SSVGI_ORACLE_H

#include "../include/lgcu.h"

#ifdef __cplusplus
extern "C" {
#endif

int orc_num_threads(void);
void orc_set_num_threads(int n);

int orc_gbuffer_resolve(const lgcu_gbuffer_builder_data *params, const lgcu_draw_call_data *objects, uint32_t nObjects,
                        const lgcu_fragment *fragments, uint64_t fragmentPitchBytes, const lgcu_clear_values *clear,
                        const lgcu_image *albedo, const lgcu_image *emissive, const lgcu_image *normal,
                        const lgcu_image *depthMoments, const lgcu_image *depthStencil, const lgcu_rows *rows);
int orc_direct_light(const lgcu_direct_lighting_data *params, const lgcu_image *albedo, const lgcu_image *emissive,
                     const lgcu_image *normal, const lgcu_image *depthStencil, const lgcu_image *shadowMap,
                     const lgcu_image *directLight, const lgcu_rows *rows);
int orc_mip_level(const lgcu_mip_level_builder_data *params, const lgcu_image *srcLevel, const lgcu_image *dstLevel,
                  const lgcu_rows *rows);
int orc_blur_level(const lgcu_blur_layer_builder_data *params, const lgcu_image *srcLevel, const lgcu_image *dstLevel,
                   const lgcu_rows *rows);
int orc_gi_gather(const lgcu_indirect_lighting_data *params, const lgcu_image *blurredDirectLight,
                  const lgcu_image *blurredDepthMoments, const lgcu_image *normal, const lgcu_image *depthStencil,
                  const lgcu_image *indirectLight, uint32_t flags, const lgcu_rows *rows);
/* work counters (pixels, march samples, horizon hits) of the last orc_gi_gather call whose flags had bit 31 set */
void orc_gi_gather_counters(unsigned long long out[3]);
int orc_denoise(const lgcu_denoiser_data *params, const lgcu_image *noisy, const lgcu_image *normal,
                const lgcu_image *depthMoments, const lgcu_image *denoised, const lgcu_rows *rows);
int orc_final_gather(const lgcu_final_gatherer_data *params, const lgcu_image *directLight,
                     const lgcu_image *blurredDirectLight, const lgcu_image *albedo, const lgcu_image *indirectLight,
                     const lgcu_image *swapchain, const lgcu_rows *rows);

int orc_deinterleave(const lgcu_interleave_data *params, const lgcu_image *interleaved, const lgcu_image *deinterleaved, const lgcu_rows *rows);
int orc_interleave(const lgcu_interleave_data *params, const lgcu_image *deinterleaved, const lgcu_image *interleaved, const lgcu_rows *rows);
int orc_debug_overlay(const lgcu_debug_quad_data *params, const lgcu_image *src, const lgcu_image *target, const lgcu_rows *rows);

#ifdef __cplusplus
}
#endif
#endif
