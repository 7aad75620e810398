/* Declarations-only stand-in for <vulkan/vulkan.h>: exactly the types, enumerants and entry points integration/vulkan/ uses, spelled
 * as the Vulkan 1.1 API publishes them (field order included), so that tests/test_integration_syntax.py can compile that code
 * in an image that has no Vulkan SDK.  Values of the enumerants are placeholders — NOTHING here may be linked or run; a real build
 * uses the real header. */
#ifndef F184_VK_STUB_H
#define F184_VK_STUB_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#define VK_DEFINE_HANDLE(n) typedef struct n##_T* n;
VK_DEFINE_HANDLE(VkInstance) VK_DEFINE_HANDLE(VkPhysicalDevice) VK_DEFINE_HANDLE(VkDevice) VK_DEFINE_HANDLE(VkCommandBuffer)
VK_DEFINE_HANDLE(VkImage) VK_DEFINE_HANDLE(VkBuffer) VK_DEFINE_HANDLE(VkDeviceMemory) VK_DEFINE_HANDLE(VkSemaphore)
#define VK_NULL_HANDLE 0
#define VK_WHOLE_SIZE (~0ULL)
#define VK_QUEUE_FAMILY_EXTERNAL (~1U)
#define VK_MAX_MEMORY_TYPES 32
#define VK_MAX_MEMORY_HEAPS 16
typedef uint64_t VkDeviceSize;
typedef uint32_t VkFlags;
typedef uint32_t VkBool32;
typedef VkFlags VkImageUsageFlags, VkBufferUsageFlags, VkPipelineStageFlags, VkAccessFlags, VkMemoryPropertyFlags, VkImageAspectFlags, VkDependencyFlags,
    VkExternalMemoryHandleTypeFlags, VkExternalSemaphoreHandleTypeFlags, VkImageCreateFlags, VkBufferCreateFlags, VkSemaphoreCreateFlags, VkMemoryHeapFlags;
typedef enum VkResult { VK_SUCCESS = 0, VK_ERROR_FEATURE_NOT_PRESENT = -8, VK_ERROR_EXTENSION_NOT_PRESENT = -7 } VkResult;
typedef enum VkStructureType {
    VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO, VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO, VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO, VK_STRUCTURE_TYPE_SEMAPHORE_CREATE_INFO,
    VK_STRUCTURE_TYPE_BUFFER_MEMORY_BARRIER, VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER, VK_STRUCTURE_TYPE_MEMORY_DEDICATED_ALLOCATE_INFO,
    VK_STRUCTURE_TYPE_EXPORT_MEMORY_ALLOCATE_INFO, VK_STRUCTURE_TYPE_EXTERNAL_MEMORY_IMAGE_CREATE_INFO, VK_STRUCTURE_TYPE_EXTERNAL_MEMORY_BUFFER_CREATE_INFO,
    VK_STRUCTURE_TYPE_EXPORT_SEMAPHORE_CREATE_INFO, VK_STRUCTURE_TYPE_MEMORY_GET_FD_INFO_KHR, VK_STRUCTURE_TYPE_SEMAPHORE_GET_FD_INFO_KHR
} VkStructureType;
typedef enum VkFormat {
    VK_FORMAT_UNDEFINED, VK_FORMAT_R8G8B8A8_UNORM, VK_FORMAT_R8G8B8A8_SNORM, VK_FORMAT_R16G16_UINT, VK_FORMAT_R16G16B16A16_UNORM, VK_FORMAT_R16G16B16A16_SFLOAT,
    VK_FORMAT_R32_UINT, VK_FORMAT_R32_SFLOAT, VK_FORMAT_R32G32B32A32_SFLOAT, VK_FORMAT_D32_SFLOAT_S8_UINT
} VkFormat;
typedef enum VkImageType { VK_IMAGE_TYPE_2D = 1 } VkImageType;
typedef enum VkImageTiling { VK_IMAGE_TILING_OPTIMAL, VK_IMAGE_TILING_LINEAR } VkImageTiling;
typedef enum VkSharingMode { VK_SHARING_MODE_EXCLUSIVE } VkSharingMode;
typedef enum VkSampleCountFlagBits { VK_SAMPLE_COUNT_1_BIT = 1 } VkSampleCountFlagBits;
typedef enum VkImageLayout {
    VK_IMAGE_LAYOUT_UNDEFINED, VK_IMAGE_LAYOUT_GENERAL, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL,
    VK_IMAGE_LAYOUT_DEPTH_STENCIL_READ_ONLY_OPTIMAL
} VkImageLayout;
enum { VK_IMAGE_ASPECT_COLOR_BIT = 1, VK_IMAGE_ASPECT_DEPTH_BIT = 2 };
enum { VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT = 1 };
enum { VK_BUFFER_USAGE_TRANSFER_DST_BIT = 2 };
enum { VK_IMAGE_USAGE_TRANSFER_SRC_BIT = 1, VK_IMAGE_USAGE_SAMPLED_BIT = 4, VK_IMAGE_USAGE_STORAGE_BIT = 8, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT = 16 };
enum { VK_PIPELINE_STAGE_TOP_OF_PIPE_BIT = 1, VK_PIPELINE_STAGE_FRAGMENT_SHADER_BIT = 0x80, VK_PIPELINE_STAGE_COLOR_ATTACHMENT_OUTPUT_BIT = 0x400,
       VK_PIPELINE_STAGE_TRANSFER_BIT = 0x1000, VK_PIPELINE_STAGE_BOTTOM_OF_PIPE_BIT = 0x2000 };
enum { VK_ACCESS_SHADER_READ_BIT = 0x20, VK_ACCESS_COLOR_ATTACHMENT_WRITE_BIT = 0x100, VK_ACCESS_TRANSFER_WRITE_BIT = 0x1000 };
enum { VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT = 1 };
enum { VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT = 1 };
typedef struct VkExtent3D { uint32_t width, height, depth; } VkExtent3D;
typedef struct VkOffset3D { int32_t x, y, z; } VkOffset3D;
typedef struct VkMemoryRequirements { VkDeviceSize size, alignment; uint32_t memoryTypeBits; } VkMemoryRequirements;
typedef struct VkMemoryType { VkMemoryPropertyFlags propertyFlags; uint32_t heapIndex; } VkMemoryType;
typedef struct VkMemoryHeap { VkDeviceSize size; VkMemoryHeapFlags flags; } VkMemoryHeap;
typedef struct VkPhysicalDeviceMemoryProperties {
    uint32_t memoryTypeCount; VkMemoryType memoryTypes[VK_MAX_MEMORY_TYPES]; uint32_t memoryHeapCount; VkMemoryHeap memoryHeaps[VK_MAX_MEMORY_HEAPS];
} VkPhysicalDeviceMemoryProperties;
typedef struct VkMemoryAllocateInfo { VkStructureType sType; const void* pNext; VkDeviceSize allocationSize; uint32_t memoryTypeIndex; } VkMemoryAllocateInfo;
typedef struct VkMemoryDedicatedAllocateInfo { VkStructureType sType; const void* pNext; VkImage image; VkBuffer buffer; } VkMemoryDedicatedAllocateInfo;
typedef struct VkExportMemoryAllocateInfo { VkStructureType sType; const void* pNext; VkExternalMemoryHandleTypeFlags handleTypes; } VkExportMemoryAllocateInfo;
typedef struct VkExternalMemoryImageCreateInfo { VkStructureType sType; const void* pNext; VkExternalMemoryHandleTypeFlags handleTypes; } VkExternalMemoryImageCreateInfo;
typedef struct VkExternalMemoryBufferCreateInfo { VkStructureType sType; const void* pNext; VkExternalMemoryHandleTypeFlags handleTypes; } VkExternalMemoryBufferCreateInfo;
typedef struct VkExportSemaphoreCreateInfo { VkStructureType sType; const void* pNext; VkExternalSemaphoreHandleTypeFlags handleTypes; } VkExportSemaphoreCreateInfo;
typedef struct VkMemoryGetFdInfoKHR { VkStructureType sType; const void* pNext; VkDeviceMemory memory; uint32_t handleType; } VkMemoryGetFdInfoKHR;
typedef struct VkSemaphoreGetFdInfoKHR { VkStructureType sType; const void* pNext; VkSemaphore semaphore; uint32_t handleType; } VkSemaphoreGetFdInfoKHR;
typedef struct VkImageCreateInfo {
    VkStructureType sType; const void* pNext; VkImageCreateFlags flags; VkImageType imageType; VkFormat format; VkExtent3D extent; uint32_t mipLevels, arrayLayers;
    VkSampleCountFlagBits samples; VkImageTiling tiling; VkImageUsageFlags usage; VkSharingMode sharingMode; uint32_t queueFamilyIndexCount;
    const uint32_t* pQueueFamilyIndices; VkImageLayout initialLayout;
} VkImageCreateInfo;
typedef struct VkBufferCreateInfo {
    VkStructureType sType; const void* pNext; VkBufferCreateFlags flags; VkDeviceSize size; VkBufferUsageFlags usage; VkSharingMode sharingMode;
    uint32_t queueFamilyIndexCount; const uint32_t* pQueueFamilyIndices;
} VkBufferCreateInfo;
typedef struct VkSemaphoreCreateInfo { VkStructureType sType; const void* pNext; VkSemaphoreCreateFlags flags; } VkSemaphoreCreateInfo;
typedef struct VkImageSubresource { VkImageAspectFlags aspectMask; uint32_t mipLevel, arrayLayer; } VkImageSubresource;
typedef struct VkSubresourceLayout { VkDeviceSize offset, size, rowPitch, arrayPitch, depthPitch; } VkSubresourceLayout;
typedef struct VkImageSubresourceLayers { VkImageAspectFlags aspectMask; uint32_t mipLevel, baseArrayLayer, layerCount; } VkImageSubresourceLayers;
typedef struct VkImageSubresourceRange { VkImageAspectFlags aspectMask; uint32_t baseMipLevel, levelCount, baseArrayLayer, layerCount; } VkImageSubresourceRange;
typedef struct VkBufferImageCopy {
    VkDeviceSize bufferOffset; uint32_t bufferRowLength, bufferImageHeight; VkImageSubresourceLayers imageSubresource; VkOffset3D imageOffset; VkExtent3D imageExtent;
} VkBufferImageCopy;
typedef struct VkMemoryBarrier { VkStructureType sType; const void* pNext; VkAccessFlags srcAccessMask, dstAccessMask; } VkMemoryBarrier;
typedef struct VkBufferMemoryBarrier {
    VkStructureType sType; const void* pNext; VkAccessFlags srcAccessMask, dstAccessMask; uint32_t srcQueueFamilyIndex, dstQueueFamilyIndex; VkBuffer buffer;
    VkDeviceSize offset, size;
} VkBufferMemoryBarrier;
typedef struct VkImageMemoryBarrier {
    VkStructureType sType; const void* pNext; VkAccessFlags srcAccessMask, dstAccessMask; VkImageLayout oldLayout, newLayout;
    uint32_t srcQueueFamilyIndex, dstQueueFamilyIndex; VkImage image; VkImageSubresourceRange subresourceRange;
} VkImageMemoryBarrier;
typedef struct VkAllocationCallbacks VkAllocationCallbacks;
typedef void (*PFN_vkVoidFunction)(void);
typedef VkResult (*PFN_vkGetMemoryFdKHR)(VkDevice, const VkMemoryGetFdInfoKHR*, int*);
typedef VkResult (*PFN_vkGetSemaphoreFdKHR)(VkDevice, const VkSemaphoreGetFdInfoKHR*, int*);
PFN_vkVoidFunction vkGetDeviceProcAddr(VkDevice, const char*);
void vkGetPhysicalDeviceMemoryProperties(VkPhysicalDevice, VkPhysicalDeviceMemoryProperties*);
VkResult vkAllocateMemory(VkDevice, const VkMemoryAllocateInfo*, const VkAllocationCallbacks*, VkDeviceMemory*);
void vkFreeMemory(VkDevice, VkDeviceMemory, const VkAllocationCallbacks*);
VkResult vkCreateImage(VkDevice, const VkImageCreateInfo*, const VkAllocationCallbacks*, VkImage*);
void vkDestroyImage(VkDevice, VkImage, const VkAllocationCallbacks*);
VkResult vkCreateBuffer(VkDevice, const VkBufferCreateInfo*, const VkAllocationCallbacks*, VkBuffer*);
void vkDestroyBuffer(VkDevice, VkBuffer, const VkAllocationCallbacks*);
VkResult vkCreateSemaphore(VkDevice, const VkSemaphoreCreateInfo*, const VkAllocationCallbacks*, VkSemaphore*);
void vkDestroySemaphore(VkDevice, VkSemaphore, const VkAllocationCallbacks*);
void vkGetImageMemoryRequirements(VkDevice, VkImage, VkMemoryRequirements*);
void vkGetBufferMemoryRequirements(VkDevice, VkBuffer, VkMemoryRequirements*);
VkResult vkBindImageMemory(VkDevice, VkImage, VkDeviceMemory, VkDeviceSize);
VkResult vkBindBufferMemory(VkDevice, VkBuffer, VkDeviceMemory, VkDeviceSize);
void vkGetImageSubresourceLayout(VkDevice, VkImage, const VkImageSubresource*, VkSubresourceLayout*);
void vkCmdCopyImageToBuffer(VkCommandBuffer, VkImage, VkImageLayout, VkBuffer, uint32_t, const VkBufferImageCopy*);
void vkCmdPipelineBarrier(VkCommandBuffer, VkPipelineStageFlags, VkPipelineStageFlags, VkDependencyFlags, uint32_t, const VkMemoryBarrier*, uint32_t,
                          const VkBufferMemoryBarrier*, uint32_t, const VkImageMemoryBarrier*);
#ifdef __cplusplus
}
#endif
#endif
