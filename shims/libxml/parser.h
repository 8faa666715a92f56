/*
 * Minimal stand-in for <libxml/parser.h> (libxml2 is not installed in this image).
 *
 * Build-environment shim, NOT part of the render hot path: it lets the unmodified
 * reference host sources (parser.c, palettes.c, flam3.c, flam3-genome.c ...) and any
 * client of flam3.h compile and run. It declares exactly the subset of the libxml2
 * tree API that flam3 3.1.1 uses; the implementation is
 * shims/libxml/xmlshim.c, a small non-validating DOM builder.
 */
#ifndef F3B_SHIM_LIBXML_PARSER_H
#define F3B_SHIM_LIBXML_PARSER_H

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char xmlChar;

typedef enum {
   XML_ELEMENT_NODE = 1,
   XML_ATTRIBUTE_NODE = 2,
   XML_TEXT_NODE = 3,
   XML_COMMENT_NODE = 8,
   XML_DOCUMENT_NODE = 9
} xmlElementType;

#define XML_PARSE_NONET (1 << 11)

struct _xmlDoc;
struct _xmlAttr;

typedef struct _xmlNode {
   void *_private;
   xmlElementType type;
   const xmlChar *name;
   struct _xmlNode *children;
   struct _xmlNode *last;
   struct _xmlNode *parent;
   struct _xmlNode *next;
   struct _xmlNode *prev;
   struct _xmlDoc *doc;
   xmlChar *content;              /* text / comment nodes */
   struct _xmlAttr *properties;   /* element nodes */
} xmlNode;
typedef xmlNode *xmlNodePtr;

typedef struct _xmlAttr {
   void *_private;
   xmlElementType type;
   const xmlChar *name;
   struct _xmlNode *children;     /* one text node holding the value */
   struct _xmlNode *last;
   struct _xmlNode *parent;
   struct _xmlAttr *next;
   struct _xmlAttr *prev;
   struct _xmlDoc *doc;
} xmlAttr;
typedef xmlAttr *xmlAttrPtr;

typedef struct _xmlDoc {
   void *_private;
   xmlElementType type;
   char *name;
   struct _xmlNode *children;
   struct _xmlNode *last;
   xmlChar *version;
} xmlDoc;
typedef xmlDoc *xmlDocPtr;

typedef struct _xmlNs xmlNs;
typedef xmlNs *xmlNsPtr;

typedef void (*xmlFreeFunc)(void *mem);
extern xmlFreeFunc xmlFree;

xmlDocPtr xmlReadMemory(const char *buffer, int size, const char *URL,
                        const char *encoding, int options);
xmlNodePtr xmlDocGetRootElement(xmlDocPtr doc);
xmlNodePtr xmlDocSetRootElement(xmlDocPtr doc, xmlNodePtr root);
xmlDocPtr xmlNewDoc(const xmlChar *version);
void xmlFreeDoc(xmlDocPtr doc);
void xmlCleanupParser(void);

xmlChar *xmlGetProp(xmlNodePtr node, const xmlChar *name);
xmlAttrPtr xmlNewProp(xmlNodePtr node, const xmlChar *name, const xmlChar *value);
xmlChar *xmlNodeGetContent(xmlNodePtr cur);
int xmlIsBlankNode(xmlNodePtr node);

xmlNodePtr xmlNewNode(xmlNsPtr ns, const xmlChar *name);
xmlNodePtr xmlNewChild(xmlNodePtr parent, xmlNsPtr ns, const xmlChar *name,
                       const xmlChar *content);
xmlNodePtr xmlAddChild(xmlNodePtr parent, xmlNodePtr cur);
xmlNodePtr xmlCopyNode(xmlNodePtr node, int recursive);

int xmlStrcmp(const xmlChar *str1, const xmlChar *str2);

#ifdef __cplusplus
}
#endif
#endif
