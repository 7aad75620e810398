// f184_vk_interop.cpp — see f184_vk_interop.h.  Vulkan 1.1 core + VK_KHR_external_memory_fd + VK_KHR_external_semaphore_fd.
#include "f184_vk_interop.h"

#include <string.h>
#include <unistd.h>

extern "C" {

const char* const kF184VkDeviceExtensions[] = {
    "VK_KHR_external_memory", "VK_KHR_external_memory_fd", "VK_KHR_external_semaphore", "VK_KHR_external_semaphore_fd", "VK_KHR_dedicated_allocation",
    "VK_KHR_get_memory_requirements2"};
const uint32_t kF184VkDeviceExtensionCount = sizeof(kF184VkDeviceExtensions) / sizeof(kF184VkDeviceExtensions[0]);

uint32_t f184vk_format(VkFormat f)
{
    switch (f)
    {
    case VK_FORMAT_R32_SFLOAT: return F184_FMT_R32_SFLOAT;
    case VK_FORMAT_R16G16B16A16_UNORM: return F184_FMT_R16G16B16A16_UNORM;
    case VK_FORMAT_R8G8B8A8_UNORM: return F184_FMT_R8G8B8A8_UNORM;
    case VK_FORMAT_R16G16B16A16_SFLOAT: return F184_FMT_R16G16B16A16_SFLOAT;
    case VK_FORMAT_R16G16_UINT: return F184_FMT_R16G16_UINT;
    case VK_FORMAT_R32G32B32A32_SFLOAT: return F184_FMT_R32G32B32A32_SFLOAT;
    case VK_FORMAT_R8G8B8A8_SNORM: return F184_FMT_R8G8B8A8_SNORM;
    case VK_FORMAT_R32_UINT: return F184_FMT_R32_UINT;
    default: return F184_FMT_UNDEFINED;
    }
}

static uint32_t device_local_type(VkPhysicalDevice phys, uint32_t type_bits)
{
    VkPhysicalDeviceMemoryProperties mp;
    vkGetPhysicalDeviceMemoryProperties(phys, &mp);
    for (uint32_t i = 0; i < mp.memoryTypeCount; i++)
        if ((type_bits & (1u << i)) && (mp.memoryTypes[i].propertyFlags & VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT)) return i;
    return UINT32_MAX;
}

// a dedicated allocation for `image` or `buffer`, exportable as an opaque fd, and the fd itself
static VkResult allocate_exported(VkPhysicalDevice phys, VkDevice dev, const VkMemoryRequirements& req, VkImage image, VkBuffer buffer, F184VkShared* out)
{
    const uint32_t type = device_local_type(phys, req.memoryTypeBits);
    if (type == UINT32_MAX) return VK_ERROR_FEATURE_NOT_PRESENT;
    VkMemoryDedicatedAllocateInfo dedicated = {VK_STRUCTURE_TYPE_MEMORY_DEDICATED_ALLOCATE_INFO};
    dedicated.image = image;
    dedicated.buffer = buffer;
    VkExportMemoryAllocateInfo exported = {VK_STRUCTURE_TYPE_EXPORT_MEMORY_ALLOCATE_INFO};
    exported.pNext = &dedicated;
    exported.handleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
    VkMemoryAllocateInfo info = {VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO};
    info.pNext = &exported;
    info.allocationSize = req.size;
    info.memoryTypeIndex = type;
    VkResult r = vkAllocateMemory(dev, &info, nullptr, &out->memory);
    if (r != VK_SUCCESS) return r;
    out->alloc_size = req.size;
    PFN_vkGetMemoryFdKHR get_fd = (PFN_vkGetMemoryFdKHR)vkGetDeviceProcAddr(dev, "vkGetMemoryFdKHR");
    if (!get_fd) return VK_ERROR_EXTENSION_NOT_PRESENT;
    VkMemoryGetFdInfoKHR fd_info = {VK_STRUCTURE_TYPE_MEMORY_GET_FD_INFO_KHR};
    fd_info.memory = out->memory;
    fd_info.handleType = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
    return get_fd(dev, &fd_info, &out->fd);
}

VkResult f184vk_create_image(VkPhysicalDevice phys, VkDevice dev, VkFormat format, uint32_t width, uint32_t height, VkImageUsageFlags usage, F184VkShared* out)
{
    memset(out, 0, sizeof(*out));
    out->fd = -1;
    out->format = format; out->width = width; out->height = height;
    // libf184's slots are pitch-linear device memory (f184_image_desc): LINEAR tiling, whose layout Vulkan reports below
    VkExternalMemoryImageCreateInfo external = {VK_STRUCTURE_TYPE_EXTERNAL_MEMORY_IMAGE_CREATE_INFO};
    external.handleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
    VkImageCreateInfo info = {VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO};
    info.pNext = &external;
    info.imageType = VK_IMAGE_TYPE_2D;
    info.format = format;
    info.extent.width = width; info.extent.height = height; info.extent.depth = 1;
    info.mipLevels = 1; info.arrayLayers = 1;
    info.samples = VK_SAMPLE_COUNT_1_BIT;
    info.tiling = VK_IMAGE_TILING_LINEAR;
    info.usage = usage;
    info.sharingMode = VK_SHARING_MODE_EXCLUSIVE;
    info.initialLayout = VK_IMAGE_LAYOUT_UNDEFINED;
    VkResult r = vkCreateImage(dev, &info, nullptr, &out->image);
    if (r != VK_SUCCESS) return r;
    VkMemoryRequirements req;
    vkGetImageMemoryRequirements(dev, out->image, &req);
    r = allocate_exported(phys, dev, req, out->image, VK_NULL_HANDLE, out);
    if (r != VK_SUCCESS) return r;
    r = vkBindImageMemory(dev, out->image, out->memory, 0);
    if (r != VK_SUCCESS) return r;
    VkImageSubresource sub = {VK_IMAGE_ASPECT_COLOR_BIT, 0, 0};
    VkSubresourceLayout layout;
    vkGetImageSubresourceLayout(dev, out->image, &sub, &layout);
    out->offset = layout.offset;
    out->row_pitch = (uint32_t)layout.rowPitch;
    return VK_SUCCESS;
}

VkResult f184vk_create_depth_copy(VkPhysicalDevice phys, VkDevice dev, uint32_t width, uint32_t height, F184VkShared* out)
{
    memset(out, 0, sizeof(*out));
    out->fd = -1;
    out->format = VK_FORMAT_R32_SFLOAT; out->width = width; out->height = height;
    out->row_pitch = width * 4u;                 // vkCmdCopyImageToBuffer with bufferRowLength 0 packs the rows tightly
    VkExternalMemoryBufferCreateInfo external = {VK_STRUCTURE_TYPE_EXTERNAL_MEMORY_BUFFER_CREATE_INFO};
    external.handleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;
    VkBufferCreateInfo info = {VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO};
    info.pNext = &external;
    info.size = (VkDeviceSize)width * height * 4u;
    info.usage = VK_BUFFER_USAGE_TRANSFER_DST_BIT;
    info.sharingMode = VK_SHARING_MODE_EXCLUSIVE;
    VkResult r = vkCreateBuffer(dev, &info, nullptr, &out->buffer);
    if (r != VK_SUCCESS) return r;
    VkMemoryRequirements req;
    vkGetBufferMemoryRequirements(dev, out->buffer, &req);
    r = allocate_exported(phys, dev, req, VK_NULL_HANDLE, out->buffer, out);
    if (r != VK_SUCCESS) return r;
    return vkBindBufferMemory(dev, out->buffer, out->memory, 0);
}

VkResult f184vk_create_semaphore(VkDevice dev, F184VkSemaphore* out)
{
    out->semaphore = VK_NULL_HANDLE; out->fd = -1;
    VkExportSemaphoreCreateInfo exported = {VK_STRUCTURE_TYPE_EXPORT_SEMAPHORE_CREATE_INFO};
    exported.handleTypes = VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT;
    VkSemaphoreCreateInfo info = {VK_STRUCTURE_TYPE_SEMAPHORE_CREATE_INFO};
    info.pNext = &exported;
    VkResult r = vkCreateSemaphore(dev, &info, nullptr, &out->semaphore);
    if (r != VK_SUCCESS) return r;
    PFN_vkGetSemaphoreFdKHR get_fd = (PFN_vkGetSemaphoreFdKHR)vkGetDeviceProcAddr(dev, "vkGetSemaphoreFdKHR");
    if (!get_fd) return VK_ERROR_EXTENSION_NOT_PRESENT;
    VkSemaphoreGetFdInfoKHR fd_info = {VK_STRUCTURE_TYPE_SEMAPHORE_GET_FD_INFO_KHR};
    fd_info.semaphore = out->semaphore;
    fd_info.handleType = VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT;
    return get_fd(dev, &fd_info, &out->fd);
}

void f184vk_destroy(VkDevice dev, F184VkShared* s)
{
    if (s->fd >= 0) close(s->fd);                // never imported: the descriptor is still ours
    if (s->image) vkDestroyImage(dev, s->image, nullptr);
    if (s->buffer) vkDestroyBuffer(dev, s->buffer, nullptr);
    if (s->memory) vkFreeMemory(dev, s->memory, nullptr);
    memset(s, 0, sizeof(*s));
    s->fd = -1;
}

void f184vk_destroy_semaphore(VkDevice dev, F184VkSemaphore* s)
{
    if (s->fd >= 0) close(s->fd);
    if (s->semaphore) vkDestroySemaphore(dev, s->semaphore, nullptr);
    s->semaphore = VK_NULL_HANDLE; s->fd = -1;
}

int f184vk_import(f184_ctx* ctx, uint32_t slot, F184VkShared* s)
{
    f184_image_desc d;
    memset(&d, 0, sizeof(d));
    d.format = f184vk_format(s->format);
    if (d.format == F184_FMT_UNDEFINED || s->fd < 0) return F184_ERR_INVALID_ARGUMENT;
    d.width = s->width; d.height = s->height; d.depth = 1;
    d.row_pitch_bytes = s->row_pitch;
    const int rc = f184_import_external_memory_fd(ctx, slot, s->fd, (uint64_t)s->alloc_size, (uint64_t)s->offset, &d);
    if (rc == F184_OK) s->fd = -1;               // cudaImportExternalMemory owns the descriptor now
    return rc;
}

int f184vk_import_semaphores(f184_ctx* ctx, F184VkSemaphore* gbuffer_done, F184VkSemaphore* gi_done)
{
    const int rc = f184_import_semaphores_fd(ctx, gbuffer_done->fd, gi_done->fd);
    if (rc == F184_OK) gbuffer_done->fd = gi_done->fd = -1;
    return rc;
}

void f184vk_cmd_copy_depth(VkCommandBuffer cmd, VkImage depth_stencil, VkImageLayout layout, const F184VkShared* dst)
{
    // the depth aspect of D32_SFLOAT_S8_UINT is copied as tightly packed 32-bit floats (Vulkan spec, "Copying Data Between Buffers and Images")
    VkBufferImageCopy region;
    memset(&region, 0, sizeof(region));
    region.imageSubresource.aspectMask = VK_IMAGE_ASPECT_DEPTH_BIT;
    region.imageSubresource.layerCount = 1;
    region.imageExtent.width = dst->width; region.imageExtent.height = dst->height; region.imageExtent.depth = 1;
    vkCmdCopyImageToBuffer(cmd, depth_stencil, layout, dst->buffer, 1, &region);
}

static void ownership_barrier(VkCommandBuffer cmd, const F184VkShared* s, uint32_t src_family, uint32_t dst_family, VkImageLayout layout,
                              VkPipelineStageFlags src_stage, VkAccessFlags src_access, VkPipelineStageFlags dst_stage, VkAccessFlags dst_access)
{
    if (s->image)
    {
        VkImageMemoryBarrier b = {VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER};
        b.srcAccessMask = src_access; b.dstAccessMask = dst_access;
        b.oldLayout = layout; b.newLayout = layout;            // CUDA sees bytes, not layouts: LINEAR images keep VK_IMAGE_LAYOUT_GENERAL across the hand-over
        b.srcQueueFamilyIndex = src_family; b.dstQueueFamilyIndex = dst_family;
        b.image = s->image;
        b.subresourceRange.aspectMask = VK_IMAGE_ASPECT_COLOR_BIT;
        b.subresourceRange.levelCount = 1; b.subresourceRange.layerCount = 1;
        vkCmdPipelineBarrier(cmd, src_stage, dst_stage, 0, 0, nullptr, 0, nullptr, 1, &b);
    }
    else
    {
        VkBufferMemoryBarrier b = {VK_STRUCTURE_TYPE_BUFFER_MEMORY_BARRIER};
        b.srcAccessMask = src_access; b.dstAccessMask = dst_access;
        b.srcQueueFamilyIndex = src_family; b.dstQueueFamilyIndex = dst_family;
        b.buffer = s->buffer; b.offset = 0; b.size = VK_WHOLE_SIZE;
        vkCmdPipelineBarrier(cmd, src_stage, dst_stage, 0, 0, nullptr, 1, &b, 0, nullptr);
    }
}

void f184vk_cmd_release(VkCommandBuffer cmd, const F184VkShared* s, uint32_t queue_family, VkImageLayout layout, VkPipelineStageFlags src_stage, VkAccessFlags src_access)
{
    ownership_barrier(cmd, s, queue_family, VK_QUEUE_FAMILY_EXTERNAL, layout, src_stage, src_access, VK_PIPELINE_STAGE_BOTTOM_OF_PIPE_BIT, 0);
}

void f184vk_cmd_acquire(VkCommandBuffer cmd, const F184VkShared* s, uint32_t queue_family, VkImageLayout layout, VkPipelineStageFlags dst_stage, VkAccessFlags dst_access)
{
    ownership_barrier(cmd, s, VK_QUEUE_FAMILY_EXTERNAL, queue_family, layout, VK_PIPELINE_STAGE_TOP_OF_PIPE_BIT, 0, dst_stage, dst_access);
}

}  // extern "C"
