// f184_vk_interop.h — the Vulkan side of the libf184 boundary, on the Vulkan C API alone.
//
// What Final184's RHI has to do so that the images the voxel-GI path reads and writes can be shared with CUDA, written as code
// instead of prose (the reference enables only VK_KHR_swapchain, RHI/Private/Vulkan/DeviceVk.cpp:266, and
// allocates every image from VMA pools, :345-386 — neither is exportable):
//   * the device extensions to add at DeviceVk.cpp:266,
//   * a LINEAR-tiled colour image on a dedicated, exportable allocation (G-buffer normals/material, the indirect and AO outputs),
//   * an exportable buffer + a depth-aspect copy for the D32_SFLOAT_S8_UINT attachments (MegaPipeline.cpp:392-394, 425-427:
//     combined depth-stencil cannot be linear, and vkCmdCopyImage cannot change aspects — vkCmdCopyImageToBuffer can),
//   * an exportable semaphore pair for the two halves of the frame (RHI/Private/Vulkan/CommandListVk.h:17-24),
//   * queue-family release / acquire barriers towards VK_QUEUE_FAMILY_EXTERNAL (the RHI's AccessTracker knows no external owner),
//   * the hand-over of all of it to libf184 (f184_import_external_memory_fd / f184_import_semaphores_fd).
// final184_rhi.patch (same directory) wires these into the reference's classes.  NOT BUILT OR RUN in this repository's image — it has
// no Vulkan headers, loader or driver; tests/test_integration_syntax.py compiles this file against a declarations-only stub so that
// at least the code is well-formed C++ against the Vulkan API as published.
#pragma once
#include <vulkan/vulkan.h>

#include "f184.h"

#ifdef __cplusplus
extern "C" {
#endif

// append to the extension list of vkCreateDevice (RHI/Private/Vulkan/DeviceVk.cpp:266)
extern const char* const kF184VkDeviceExtensions[];
extern const uint32_t kF184VkDeviceExtensionCount;

typedef struct F184VkShared {
    VkImage image;              // colour images: LINEAR tiling, one mip, one layer; VK_NULL_HANDLE for a buffer
    VkBuffer buffer;            // depth copies: tightly packed R32 rows; VK_NULL_HANDLE for an image
    VkDeviceMemory memory;      // dedicated allocation, exported once as an opaque fd
    VkDeviceSize alloc_size;
    VkDeviceSize offset;        // of texel (0, 0) inside the allocation
    uint32_t row_pitch;         // bytes
    VkFormat format;
    uint32_t width, height;
    int fd;                     // -1 once libf184 has imported it (the import takes ownership of the descriptor)
} F184VkShared;

typedef struct F184VkSemaphore {
    VkSemaphore semaphore;
    int fd;
} F184VkSemaphore;

VkResult f184vk_create_image(VkPhysicalDevice phys, VkDevice dev, VkFormat format, uint32_t width, uint32_t height, VkImageUsageFlags usage, F184VkShared* out);
VkResult f184vk_create_depth_copy(VkPhysicalDevice phys, VkDevice dev, uint32_t width, uint32_t height, F184VkShared* out);
VkResult f184vk_create_semaphore(VkDevice dev, F184VkSemaphore* out);
void f184vk_destroy(VkDevice dev, F184VkShared* s);
void f184vk_destroy_semaphore(VkDevice dev, F184VkSemaphore* s);

// hand-over to libf184: returns an F184_* status
int f184vk_import(f184_ctx* ctx, uint32_t slot, F184VkShared* s);
int f184vk_import_semaphores(f184_ctx* ctx, F184VkSemaphore* gbuffer_done, F184VkSemaphore* gi_done);

// recorded at the end of command list A / the start of list B
void f184vk_cmd_copy_depth(VkCommandBuffer cmd, VkImage depth_stencil, VkImageLayout layout, const F184VkShared* dst);
void f184vk_cmd_release(VkCommandBuffer cmd, const F184VkShared* s, uint32_t queue_family, VkImageLayout layout, VkPipelineStageFlags src_stage, VkAccessFlags src_access);
void f184vk_cmd_acquire(VkCommandBuffer cmd, const F184VkShared* s, uint32_t queue_family, VkImageLayout layout, VkPipelineStageFlags dst_stage, VkAccessFlags dst_access);

// VkFormat <-> f184_format for the formats the path's slots use (F184_FMT_UNDEFINED otherwise)
uint32_t f184vk_format(VkFormat format);

#ifdef __cplusplus
}
#endif
