// Build include/f184_renderer.hpp against the CPU oracle (TEST INFRASTRUCTURE): the oracle mirrors the C-ABI with an f184o_
// prefix, so renaming the declarations of f184.h is all it takes.
#pragma once
#define f184_create f184o_create
#define f184_destroy f184o_destroy
#define f184_last_error f184o_last_error
#define f184_scene_upload f184o_scene_upload
#define f184_texture_upload f184o_texture_upload
#define f184_material_set f184o_material_set
#define f184_upload_image f184o_upload_image
#define f184_readback f184o_readback
#define f184_image_info f184o_image_info
#define f184_voxelize f184o_voxelize
#define f184_gtao f184o_gtao
#define f184_trace_indirect f184o_trace_indirect
#define f184_blur_indirect f184o_blur_indirect
#define f184_lighting_deferred f184o_lighting_deferred
#define f184_composite f184o_composite
#define f184_copy_indirect_to_history f184o_copy_indirect_to_history
#define f184_copy_taa_to_history f184o_copy_taa_to_history
